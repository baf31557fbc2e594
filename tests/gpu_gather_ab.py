"""A/B of gather_mean builds in ONE process (interleaved rounds, so clock drift cancels):
    python tests/gpu_gather_ab.py libA.so libB.so ...
Every library is dlopen'ed on its own and called through the C ABI with the same device buffers;
the outputs must be bit-identical across the builds."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from multimodal_concept_learning_b200 import _lib  # noqa: E402

libs = []
for path in sys.argv[1:]:
    lib = C.CDLL(path)
    res, args = _lib.SIGNATURES["mcl_gather_mean"]
    lib.mcl_gather_mean.restype, lib.mcl_gather_mean.argtypes = res, args
    lib.mcl_last_error.restype = C.c_char_p
    libs.append((path.split("/")[-1], lib))
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(77)
V, Q = 152064, 65536
stream = torch.cuda.current_stream().cuda_stream
for D, lo, hi in ((3584, 1, 5), (1152, 1, 5), (3584, 1, 2), (768, 1, 9)):
    table = torch.randn(V, D, generator=g, device=dev).to(torch.bfloat16)
    lens = torch.randint(lo, hi, (Q,), generator=g, device=dev)
    offs = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), lens.cumsum(0)])
    ids = torch.randint(0, V, (int(offs[-1]),), generator=g, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    for normalize in (1, 0):
        outs, times = {}, {n: [] for n, _ in libs}

        def call(lib, out):
            rc = lib.mcl_gather_mean(table.data_ptr(), 0, V, D, table.stride(0), offs.data_ptr(), ids.data_ptr(), Q,
                                     normalize, out.data_ptr(), D, flag.data_ptr(), stream)
            assert rc == 0, lib.mcl_last_error()
        for name, lib in libs:
            outs[name] = torch.empty((Q, D), dtype=torch.bfloat16, device=dev)
            for _ in range(3):
                call(lib, outs[name])
        torch.cuda.synchronize()
        for _ in range(3):
            for name, lib in libs:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    call(lib, outs[name])
                e1.record()
                torch.cuda.synchronize()
                times[name].append(e0.elapsed_time(e1) / 10)
        ref = outs[libs[0][0]]
        same = all(torch.equal(ref.view(torch.int16), o.view(torch.int16)) for o in outs.values())
        b = ids.numel() * D * 2.0 + Q * D * 2.0 + ids.numel() * 8.0 + (Q + 1) * 8.0
        print(f"D={D} ids/row {lo}-{hi - 1} normalize={normalize} bit-identical={same}: " +
              "  ".join(f"{n} {min(t) * 1e3:.0f} us ({b / min(t) / 1e6:.0f} GB/s)" for n, t in times.items()))
    del table
