"""Measurement script: gather_mean over ids-per-row distributions and ring depths (D = 3584)."""
import sys
import torch
sys.path.insert(0, ".")
import multimodal_concept_learning_b200 as mcl  # noqa: E402
from tests.gpu_rowkernels import timed  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(1)
for V, D in [(152064, 3584), (152064, 2048)]:
    table = torch.randn(V, D, generator=g, device="cuda").to(torch.bfloat16)
    Q = 65536
    for name, lens in [("1", torch.full((Q,), 1)), ("2", torch.full((Q,), 2)), ("3", torch.full((Q,), 3)), ("4", torch.full((Q,), 4)),
                       ("[1,5)", torch.randint(1, 5, (Q,))), ("[2,4)", torch.randint(2, 4, (Q,))),
                       ("[1,5) sorted", torch.randint(1, 5, (Q,)).sort().values), ("8", torch.full((Q,), 8))]:
        lens = lens.cuda()
        offs = torch.cat([torch.zeros(1, dtype=torch.long, device="cuda"), lens.cumsum(0)])
        ids = torch.randint(0, V, (int(offs[-1]),), generator=g, device="cuda")
        b = ids.numel() * D * 2 + Q * D * 2 + ids.numel() * 8 + (Q + 1) * 8
        line = f"D={D} ids/row {name:14s} nnz={ids.numel():7d}:"
        for variant in (0, 1):
            old = mcl.set_option(17, variant)
            ms = timed(lambda: mcl.gather_mean(table, offs, ids, False, validate=False))
            msn = timed(lambda: mcl.gather_mean(table, offs, ids, True, validate=False))
            mcl.set_option(17, old)
            line += f"  v{variant} {ms * 1e3:7.1f} us {b / ms / 1e6:5.0f} GB/s (normalised {msn * 1e3:7.1f} us {b / msn / 1e6:5.0f})"
        print(line, flush=True)
    del table
