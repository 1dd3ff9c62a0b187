"""Developer tool (no GPU): mutated copies of optimised IR text ("mmir 1", what the reference-side binding hands to mmb_load_ir)
go through the loader and the CUDA emitter.  Every outcome but MathMapError / success is a bug.
Usage: python tools/fuzz_ir_loader.py SEED COUNT"""
import glob, random, sys, os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mathmap_b200 as mb
rng=random.Random(int(sys.argv[1])); N=int(sys.argv[2])
irs=[open(p).read() for p in sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'ir', '*.mmir')))]
EX=os.path.join(ROOT, 'tests', 'golden', 'filters', 'examples')
for p in sorted(glob.glob(EX+'/*/*.mm'))[::9]:
    irs.append(mb.Module.from_file(p).ir)
ok=err=emit=0
for k in range(N):
    s=rng.choice(irs)
    for _ in range(rng.randint(1,3)):
        c=rng.random()
        if c<0.3:   # delete a span
            i=rng.randrange(len(s)); j=min(len(s), i+rng.randint(1,30)); s=s[:i]+s[j:]
        elif c<0.5: # change a number
            ms=list(re.finditer(r"\d+", s))
            if ms:
                m=rng.choice(ms); s=s[:m.start()]+str(rng.choice([0,1,7,99,100000,-1,2**31]))+s[m.end():]
        elif c<0.7: # swap op name
            ms=list(re.finditer(r"\(op ([A-Za-z0-9_]+)", s))
            if ms:
                m=rng.choice(ms); m2=rng.choice(ms); s=s[:m.start(1)]+m2.group(1)+s[m.end(1):]
        elif c<0.85: # insert parens / tokens
            i=rng.randrange(len(s)); s=s[:i]+rng.choice(["(",")"," %1.1 "," i:5 "," f:1.5 ","(phi","(if","(while","(closure"])+s[i:]
        else:       # duplicate a line
            ls=s.split("\n"); i=rng.randrange(len(ls)); ls.insert(i, ls[i]); s="\n".join(ls)
    try:
        m=mb.Module(ir=s); ok+=1
        try:
            m.cuda_source; emit+=1
        except mb.MathMapError: pass
    except mb.MathMapError:
        err+=1
print("loaded", ok, "emitted", emit, "rejected", err)
