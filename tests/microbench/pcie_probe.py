"""Host-link probe for the end-to-end arm of bench.py (development aid): pinned H2D alone, H2D with a concurrent D2H at half
the volume (the arm's traffic pattern), the same from write-combined pinned memory, and by copy size. Prints JSON lines."""
import ctypes as C
import glob
import json
import os
import sys

import torch


def cudart():
    import nvidia.cuda_runtime as m  # the runtime torch ships
    path = glob.glob(os.path.join(os.path.dirname(m.__file__), "lib", "libcudart.so*"))[0]
    return C.CDLL(path)


def main():
    torch.cuda.init()
    dev = torch.device("cuda", 0)
    rt = cudart()
    GiB = 1 << 30
    d_in = torch.empty(GiB // 8, dtype=torch.int64, device=dev)
    d_out = torch.empty(GiB // 16, dtype=torch.int64, device=dev)
    h_out = torch.empty(GiB // 16, dtype=torch.int64).pin_memory()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def host_buffer(kind):
        if kind == "pinned":
            return torch.empty(GiB // 8, dtype=torch.int64).pin_memory()
        p = C.c_void_p()
        flags = 4 if kind == "write_combined" else 0  # cudaHostAllocWriteCombined
        rc = rt.cudaHostAlloc(C.byref(p), C.c_size_t(GiB), C.c_uint(flags))
        if rc != 0:
            raise RuntimeError(f"cudaHostAlloc rc={rc}")
        buf = (C.c_byte * GiB).from_address(p.value)
        t = torch.frombuffer(buf, dtype=torch.int64)
        return t

    for kind in ("pinned", "host_alloc", "write_combined"):
        h_in = host_buffer(kind)
        h_in.fill_(3)
        for chunk_mib in (1024, 64, 16):
            n = chunk_mib * (1 << 20) // 8
            for both in (False, True):
                def run():
                    for o in range(0, GiB // 8, n):
                        with torch.cuda.stream(s_in):
                            d_in[o:o + n].copy_(h_in[o:o + n], non_blocking=True)
                        if both:
                            with torch.cuda.stream(s_out):
                                h_out[o // 2:(o + n) // 2].copy_(d_out[o // 2:(o + n) // 2], non_blocking=True)
                for _ in range(2):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                reps = 5
                for _ in range(reps):
                    run()
                s_in.synchronize(); s_out.synchronize()
                e1.record()
                torch.cuda.synchronize()
                import time
                t0 = time.perf_counter()
                for _ in range(reps):
                    run()
                torch.cuda.synchronize()
                dt = (time.perf_counter() - t0) / reps
                print(json.dumps({"host_memory": kind, "chunk_mib": chunk_mib, "with_d2h": both, "h2d_gbs": GiB / dt / 1e9}), flush=True)


if __name__ == "__main__":
    main()
