"""Device time of a filter's pixel kernel without torch (ctypes on libcudart for the buffer and the events): a quick A/B of
an IR-level switch on a GPU box.

    python tools/time_kernel.py FILTER.mm WIDTH HEIGHT [name=value ...] [--launches N] [--env MMB_LOOP_CARRY=0,1]

Prints one JSON line per setting of the --env variable: {"setting": .., "ms_per_launch": .., "checksum": ..} -- the checksum
(sum of all output bytes, read back once) must not depend on the setting."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import mathmap_b200 as mb
    args = sys.argv[1:]
    launches, env_name, env_values = 10, None, [None]
    if "--launches" in args:
        i = args.index("--launches"); launches = int(args[i + 1]); del args[i:i + 2]
    if "--env" in args:
        i = args.index("--env"); env_name, vals = args[i + 1].split("="); env_values = vals.split(","); del args[i:i + 2]
    path, W, H = args[0], int(args[1]), int(args[2])
    uv = dict(a.split("=") for a in args[3:])
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rt.cudaEventRecord.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    rt.cudaEventSynchronize.argtypes = [ctypes.c_void_p]
    rt.cudaEventElapsedTime.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_void_p]
    dev = ctypes.c_void_p()
    assert rt.cudaMalloc(ctypes.byref(dev), W * H * 4) == 0
    e0, e1 = ctypes.c_void_p(), ctypes.c_void_p()
    assert rt.cudaEventCreate(ctypes.byref(e0)) == 0 and rt.cudaEventCreate(ctypes.byref(e1)) == 0
    host = np.empty((H, W, 4), np.uint8)
    for setting in env_values:
        if env_name:
            os.environ[env_name] = setting
        m = mb.Module.from_file(path)
        inv = mb.Invocation(m, W, H)
        for k, v in uv.items():
            inv.set(k, float(v) if "." in v else int(v))
        inv.init_frame(0, 0.0)
        for _ in range(3):
            inv.calc_lines_device(dev.value)
        inv.synchronize()
        assert rt.cudaEventRecord(e0, None) == 0
        for _ in range(launches):
            inv.calc_lines_device(dev.value)
        assert rt.cudaEventRecord(e1, None) == 0
        assert rt.cudaEventSynchronize(e1) == 0
        ms = ctypes.c_float()
        assert rt.cudaEventElapsedTime(ctypes.byref(ms), e0, e1) == 0
        assert rt.cudaMemcpy(host.ctypes.data, dev, W * H * 4, 2) == 0
        print(json.dumps({"filter": os.path.basename(path), "size": [W, H], "setting": "%s=%s" % (env_name, setting) if env_name else None,
                          "ms_per_launch": ms.value / launches, "launches": launches, "checksum": int(host.sum(dtype=np.uint64))}), flush=True)


if __name__ == "__main__":
    main()
