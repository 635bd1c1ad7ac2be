"""GEMM probe: times wts_gemm shapes with the kernel's debug modes (WTS_GEMM_DEBUG: 1 = TMA only, 2 = MMA only,
3 = hi*hi only) to separate operand feed from tensor-pipe time.  Results in debug modes are garbage by design."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "whisper-timestamped_b200"))


def main():
    from whisper_timestamped.engine import CudaEngine
    from whisper_timestamped.model import SB16
    dev = torch.device("cuda:0")
    eng = CudaEngine.__new__(CudaEngine)
    eng.dev, eng.backend, eng.launches = dev, 0, 0
    shapes = [(24000, 5120, 1280, "enc fc1"), (24000, 1280, 5120, "enc fc2"), (24000, 1280, 1280, "enc out"),
              (128, 1280, 1280, "dec out"), (128, 5120, 1280, "dec fc1"), (128, 1280, 5120, "dec fc2")]
    for (M, N, K, name) in shapes:
        a, b = SB16(M, K, dev), SB16(N, K, dev)
        a.t.normal_()
        b.t.normal_()
        out = SB16(M, N, dev)
        x = torch.zeros(M, N, device=dev)
        for mode in (0, 1, 2, 3):
            os.environ["WTS_GEMM_DEBUG"] = str(mode)
            def run():
                if M <= 128:
                    eng.gemm(a, b, M, N, K, residual=x, ldr=N, out_f32=x, ldc=N)
                else:
                    eng.gemm(a, b, M, N, K, out_sb=out)
            for _ in range(3):
                run()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 20
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            ktiles = ((M + 127) // 128) * ((N + 127) // 128) * ((K + 63) // 64)
            per = ms * 1e-3 * 1.965e9 * 148 / ktiles
            print(f"{name:8s} M={M} N={N} K={K} mode={mode}: {ms * 1e3:9.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TF/s(alg)  "
                  f"{per:7.0f} clk per 128x128x64 k-tile per SM", flush=True)
    os.environ["WTS_GEMM_DEBUG"] = "0"


if __name__ == "__main__":
    main()
