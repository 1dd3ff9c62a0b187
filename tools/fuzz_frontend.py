"""Developer tool (no GPU): mutated copies of the example filters go through the front end, a fifth of those that still parse
through the CUDA emitter.  Every outcome but MathMapError / success is a bug (a crash would take the host application down).
Usage: python tools/fuzz_frontend.py SEED COUNT"""
import glob, random, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mathmap_b200 as mb
EX=os.path.join(ROOT, 'tests', 'golden', 'filters', 'examples')
rng=random.Random(int(sys.argv[1])); N=int(sys.argv[2])
srcs=[open(p, errors='replace').read() for p in sorted(glob.glob(EX+'/*/*.mm'))]
tokens=["(",")","[","]",":",";",",","+","-","*","/","%","^","=","==","<",">","if","then","else","end","while","do","for","..","filter","image","float","int","in","xy","ra","rgba","x","y","t","1","0.5","\"","#","::","&&","||","!"]
ok=err=0
for k in range(N):
    s=rng.choice(srcs)
    ops=rng.randint(1,4)
    for _ in range(ops):
        i=rng.randrange(len(s)+1)
        c=rng.random()
        if c<0.35:
            j=min(len(s), i+rng.randint(1,12)); s=s[:i]+s[j:]
        elif c<0.7:
            s=s[:i]+" "+rng.choice(tokens)+" "+s[i:]
        elif c<0.85:
            j=min(len(s), i+rng.randint(1,40)); s=s[:i]+s[i:j]+s[i:j]+s[j:]
        else:
            s=s[:i]+chr(rng.randrange(1,256))+s[i:]
    try:
        m=mb.Module(source=s)
        ok+=1
        if rng.random()<0.2:
            try: m.cuda_source
            except mb.MathMapError: pass
    except mb.MathMapError:
        err+=1
    except UnicodeEncodeError:
        err+=1
print("parsed", ok, "rejected", err)
