"""Measurement script (not a pytest file): the HBM-bound row kernels alone -- part (c) of the path --
with CUDA events, for ncu captures and for the roofline lines in profiles/."""
import sys

import torch

sys.path.insert(0, ".")
import multimodal_concept_learning_b200 as mcl  # noqa: E402


def timed(fn, steps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    quick = "--quick" in sys.argv
    g = torch.Generator(device="cuda").manual_seed(1)
    for V, D in ([(152064, 3584)] if quick else [(152064, 3584), (262235, 1152), (1048576, 1024)]):
        table = torch.randn(V, D, generator=g, device="cuda").to(torch.bfloat16)
        ms = timed(lambda: mcl.row_inv_norm(table))
        b = V * D * 2 + V * 4
        print(f"row_inv_norm V={V} D={D}: {ms * 1e3:.1f} us, {b / ms / 1e6:.0f} GB/s")
        for Q, lo, hi, norm in ([(65536, 1, 5, True)] if quick else
                                [(65536, 1, 5, True), (65536, 1, 5, False), (65536, 1, 2, False), (65536, 4, 5, False),
                                 (8192, 1, 5, True)]):
            lens = torch.randint(lo, hi, (Q,), generator=g, device="cuda")
            offs = torch.cat([torch.zeros(1, dtype=torch.long, device="cuda"), lens.cumsum(0)])
            ids = torch.randint(0, V, (int(offs[-1]),), generator=g, device="cuda")
            b = ids.numel() * D * 2 + Q * D * 2 + ids.numel() * 8 + (Q + 1) * 8
            for variant in (0, 1):
                old = mcl.set_option(17, variant)
                ms = timed(lambda: mcl.gather_mean(table, offs, ids, norm, validate=False))
                mcl.set_option(17, old)
                print(f"gather_mean[v{variant}] V={V} D={D} Q={Q} ids/row in [{lo},{hi}) normalize={norm}: {ms * 1e3:.1f} us, "
                      f"{b / ms / 1e6:.0f} GB/s ({b / 1e6:.0f} MB)")
        del table


if __name__ == "__main__":
    main()
