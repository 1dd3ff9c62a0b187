"""Compiles a filter's generated CUDA (as NVRTC would) with nvcc to a cubin and prints the SASS of its pixel kernel.
Usage: python tools/dump_sass.py FILTER.mm [aa=1] [precise=1] > out.sass"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mathmap_b200 as mb
m = mb.Module.from_file(sys.argv[1])
aa = int(sys.argv[2]) if len(sys.argv) > 2 else 1
precise = int(sys.argv[3]) if len(sys.argv) > 3 else 1
src = m.cuda_source.replace("#define MM_AA 0", "#define MM_AA %d" % aa).replace("#define MM_PRECISE 1", "#define MM_PRECISE %d" % precise)
d = tempfile.mkdtemp()
open(os.path.join(d, "k.cu"), "w").write(src)
rt = os.path.join(ROOT, "mathmap_b200", "csrc", "runtime")
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-cubin", "-fmad=false", "-std=c++17", "-lineinfo", "-I", rt, "-o", os.path.join(d, "k.cubin"), os.path.join(d, "k.cu")])
out = subprocess.run(["cuobjdump", "-sass", os.path.join(d, "k.cubin")], stdout=subprocess.PIPE, text=True).stdout
print(out)
