"""Extracts libnoise 1.0.0's gradient-vector table (vectortable.h, explicitly
public domain: "I'm not going to copyright a bunch of random numbers ... This
file is in the public domain") from the reference's vendored
libnoisesrc-1.0.0.zip and writes it as a plain C initialiser list shared by the
oracle (oracle/runtime/noise.c) and the device runtime
(mathmap_b200/csrc/runtime/mm_noise_table.inc).  The 256 x (x, y, z) values are
data the noise functions cannot be bit-compatible without.

Run here (needs /root/reference); the outputs are committed.
"""
import re
import sys
import zipfile

src = zipfile.ZipFile("/root/reference/libnoisesrc-1.0.0.zip").read("noise/src/vectortable.h").decode()
body = src[src.index("{", src.index("g_randomVectors")) + 1: src.rindex("}", 0, src.rindex("};") + 1)]
nums = re.findall(r"-?\d+\.?\d*(?:[eE][-+]?\d+)?", body)
assert len(nums) == 1024, len(nums)
rows = []
for i in range(256):
    x, y, z, w = nums[4 * i: 4 * i + 4]
    assert float(w) == 0.0
    rows.append("  %s, %s, %s," % (x, y, z))
text = ("/* libnoise 1.0.0 gradient vectors (public domain data, see tools/gen_noise_table.py): 256 rows of x, y, z */\n"
        + "\n".join(rows) + "\n")
for out in sys.argv[1:]:
    open(out, "w").write(text)
