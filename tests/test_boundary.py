"""CPU: the C-ABI boundary -- the library builds for sm_100a, loads, exports every symbol
include/mcl.h declares, validates arguments before touching a device, and its tile
schedule / slot map (host logic) covers every (row block, table tile) exactly once."""
import ctypes as C
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mcl.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mcl_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported(lib_built):
    from multimodal_concept_learning_b200 import _lib
    lib = C.CDLL(lib_built)
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in mcl.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in _lib.py"
    assert set(_lib.SIGNATURES) == set(syms)
    assert _lib.load().mcl_version() == 100


def test_library_is_sm100a_native(lib_built):
    out = subprocess.run(["cuobjdump", "-lelf", lib_built], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", lib_built], capture_output=True, text=True).stdout
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):      # tcgen05.mma, tcgen05.ld, TMA
        assert mnemonic in sass, f"{mnemonic} missing from SASS"
    assert "HMMA.16816" not in sass                       # no legacy mma.sync path


def test_no_cpu_fallback(lib_built):
    import multimodal_concept_learning_b200 as mcl
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mcl.concept_scan(torch.zeros(4, 8), torch.zeros(9, 8), 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mcl.row_inv_norm(torch.zeros(4, 8))
    # the product package must not import the oracle
    pkg = os.path.join(ROOT, "multimodal_concept_learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_argument_validation_needs_no_device(lib_built):
    from multimodal_concept_learning_b200 import _lib
    lib = _lib.load()
    buf = C.create_string_buffer(4096)
    p = C.addressof(buf)
    p16 = (p + 15) & ~15
    common = dict(Q=4, V=9, D=8)
    def scan(k=2, scale=1.0, dtype=0, q=p16, ldq=8):
        return lib.mcl_concept_scan(q, p16, dtype, 4, 9, 8, ldq, 8, None, None, scale, k, 0, None,
                                    p16, p16, p16, p16, 4096, None)
    assert scan(k=0) == -1 and "k=0" in _lib.last_error()
    assert scan(k=65) == -1
    assert scan(k=10) == -1                  # k > V
    assert scan(scale=0.0) == -1
    assert scan(dtype=7) == -1
    assert scan(ldq=4) == -1                 # ld < D
    assert scan(q=p16 + 2) == -2             # misaligned base
    assert scan(ldq=9, dtype=0) == -2        # pitch 18 B
    rc = scan()                              # valid arguments: fails only for lack of a device
    if not torch.cuda.is_available():
        assert rc in (-5, -3), _lib.last_error()
    with pytest.raises(_lib.MclError):
        _lib.check(scan(k=0))
    assert lib.mcl_scan_workspace_bytes(8192, 152064, 3584, 50, 0) > 0
    assert lib.mcl_sharded_gather_bytes(100, 50, 8) >= 8 * (100 * 50 * 12 + 1600)


def test_host_only_sizing_of_the_round2_entry_points(lib_built):
    """Sizing / eligibility functions that need no device: the peer-memory exchange of the sharded
    scan, the row blocking of the backward, and the argument checks in front of both."""
    from multimodal_concept_learning_b200 import _lib
    lib = _lib.load()
    # peer-memory exchange: world 2..16, Q % world == 0, (Q / world) * k % 4 == 0
    assert lib.mcl_sharded_p2p_block_bytes(8192, 50, 8) > 8192 * (12 * 50 + 16) * 2
    assert lib.mcl_sharded_p2p_block_bytes(304, 50, 2) > 0 and lib.mcl_sharded_p2p_block_bytes(304, 50, 8) > 0
    assert lib.mcl_sharded_p2p_block_bytes(301, 50, 2) == 0          # rows do not divide
    assert lib.mcl_sharded_p2p_block_bytes(304, 1, 8) == 0           # 38 rows x k = 1: pieces are not 16-byte multiples
    assert lib.mcl_sharded_p2p_block_bytes(304, 50, 1) == 0 and lib.mcl_sharded_p2p_block_bytes(3400, 50, 17) == 0
    # record of all rows + two receive areas (world slots of Q / world rows each) + result area = 4 records
    assert lib.mcl_sharded_p2p_block_bytes(8192, 50, 8) >= 4 * 8192 * (12 * 50 + 16)
    buf = C.create_string_buffer(4096)
    p16 = (C.addressof(buf) + 15) & ~15
    blocks = (C.c_void_p * 2)(p16, p16)
    def p2p(world=2, rank=0, epoch=1, Q=304, block_bytes=1 << 30, k=50):
        return lib.mcl_concept_scan_sharded_p2p(p16, p16, 0, Q, 3000, 64, 64, 64, None, None, 1.0, k, 0, None, p16, p16,
                                                p16, p16, 4096, blocks, block_bytes, world, rank, epoch, epoch, 0, None)
    assert p2p(world=1) == -1 and p2p(rank=2) == -1
    assert p2p(Q=301) == -1 and "Q % world" in _lib.last_error()
    assert p2p(block_bytes=64) == -4                                  # MCL_ERR_WORKSPACE_TOO_SMALL
    assert p2p(epoch=0) == -1 and "epoch" in _lib.last_error()
    # backward: rows per block (<= 4096, whole row blocks of 128), and the bf16-gradient precondition
    assert lib.mcl_ce_backward_block_rows(24, 262235, 0) == 128
    assert lib.mcl_ce_backward_block_rows(6750, 6000, 0) == 4096
    assert lib.mcl_ce_backward_block_rows(0, 10, 0) == 0
    def bwd(n, gt_dtype, D=64, dtype=0):
        return lib.mcl_ce_backward_ex(p16, p16, dtype, n, 6000, D, D, D, p16, p16, 1.0, 0.0, 0.0, 6000, p16, n, None,
                                      p16, gt_dtype, p16, 4096, None)
    assert bwd(6750, 0) == -1 and "one block" in _lib.last_error()    # bf16 gradient across two row blocks
    assert bwd(100, 0, dtype=1) == -1                                 # ... or from fp32 inputs
    assert bwd(100, 7) == -1 and "grad_table_dtype" in _lib.last_error()


@pytest.mark.parametrize("shape", [(16, 50257, 768), (4096, 49408, 768), (8192, 152064, 3584),
                                   (65536, 128256, 4096), (32768, 1048576, 1024), (8192, 19008, 3584),
                                   (1, 1, 8), (129, 257, 72), (700, 3000, 64), (300, 5000, 768),
                                   (8192, 76032, 3584), (1000, 262235, 1152), (32768, 1000000, 1024)])
@pytest.mark.parametrize("sm", [148, 5, 1])
@pytest.mark.parametrize("leftover", [1, 0])
def test_plan_covers_every_tile_once(lib_built, shape, sm, leftover):
    """The tile plan (csrc/plan.h) through the same functions the scan kernel and the merge
    kernel evaluate: every (row block, table tile) is scanned exactly once, no two segments
    share a slot, and the merge's slot map names exactly the slots that were written."""
    from multimodal_concept_learning_b200 import _lib
    lib = _lib.load()
    Q, V, D = shape
    old = lib.mcl_set_option(7, leftover)
    try:
        p = _lib.plan_scan(Q, V, D, sm)
        segs = _lib.plan_segments(Q, V, D, sm)
        slots_of = {}
        for rb in range(p["num_rb"]):
            a, b = C.c_int32(), C.c_int32()
            assert lib.mcl_plan_row_block_slots(Q, V, D, sm, rb, C.byref(a), C.byref(b)) == 0
            slots_of[rb] = (a.value, b.value)
    finally:
        lib.mcl_set_option(7, old)
    assert p["num_rb"] == -(-Q // 128) and p["num_vt"] == -(-V // 256) and p["num_kb"] == -(-D // 64)
    cs, S = p["cs"], p["S"]
    assert cs in (1, 2) and p["ru"] == -(-p["num_rb"] // cs)
    assert 1 <= p["grid"] <= sm and p["grid"] == p["workers"] * cs
    assert p["nslots"] == p["ru"] * cs * S * 2
    written = {}                       # slot -> (row block, column half, tiles)
    per_worker = {}
    for w, unit, vt0, vt1, j, sync in segs:
        assert 0 <= w < p["workers"] and 0 <= unit < p["ru"] and 0 <= vt0 < vt1 <= p["num_vt"]
        assert 0 <= j < S and -1 <= sync < max(1, p["nctr"])
        per_worker[w] = per_worker.get(w, 0) + (vt1 - vt0)
        for crank in range(cs):
            rb = unit * cs + crank
            for half in (0, 1):        # warps 2-5 / 6-9: columns 0-127 / 128-255 of every tile
                slot = (rb * S + j) * 2 + half
                assert slot not in written, "two segments write one slot"
                written[slot] = (rb, half, list(range(vt0, vt1)))
    assert len(per_worker) == p["workers"], "a launched worker has no work"
    for rb in range(p["num_rb"]):
        slot0, n = slots_of[rb]
        assert slot0 == rb * S * 2 and 2 <= n <= S * 2
        cover = {0: [], 1: []}
        for i in range(n):
            owner, half, t = written[slot0 + i]
            assert owner == rb and half == i % 2 and t
            cover[half] += t
        assert sorted(cover[0]) == sorted(cover[1]) == list(range(p["num_vt"])), \
            f"row block {rb}: tiles not covered exactly once"
        assert all(slot0 + i not in written for i in range(n, S * 2)), "written slot the merge skips"
    # balance of the chosen plan at full SM count
    crit = max(per_worker.values())
    if sm == 148 and p["num_rb"] * p["num_vt"] >= 148 * 8:
        ideal = p["ru"] * p["num_vt"] / (148 // cs)
        # (tail workers may carry more tiles than group members: their segments start with a
        # warm top-k filter, and the plan balances time, not tiles)
        assert crit <= 1.2 * ideal + 2, (p, crit, ideal)




def test_small_cta_counts_plan_tail_workers(lib_built):
    """tests/test_gpu_parity.py::test_tc_tail_workers_against_oracle checks tail passes and
    second-level nodes against the oracle at these CTA counts: make sure they still plan them."""
    from multimodal_concept_learning_b200 import _lib
    plans = [_lib.plan_scan(1500, 9000, 128, ctas) for ctas in (10, 14, 22, 26, 38)]
    with_tail = [p for p in plans if p["last"][0]["wr"] > 0]
    assert len(with_tail) >= 2, [p["last"] for p in plans]
    assert any(len(p["last"]) >= 2 for p in with_tail), "no plan with a second-level node"


def _check_plan(lib, _lib, Q, V, D, sm):
    """Every (row block, tile) exactly once; one writer per slot; merge-side slot map exact."""
    p = _lib.plan_scan(Q, V, D, sm)
    segs = _lib.plan_segments(Q, V, D, sm)
    cs, S, T = p["cs"], p["S"], p["num_vt"]
    assert 1 <= p["grid"] <= sm and p["workers"] * cs == p["grid"]
    covered = {}
    slots = {}
    workers = set()
    for w, unit, vt0, vt1, j, sync in segs:
        assert 0 <= w < p["workers"] and 0 <= unit < p["ru"] and 0 <= vt0 < vt1 <= T and 0 <= j < S
        assert 0 <= sync < max(1, p["nctr"])
        workers.add(w)
        assert (unit, j) not in slots, "two segments end in one slot"
        slots[(unit, j)] = (vt0, vt1)
        covered.setdefault(unit, []).append((vt0, vt1))
    assert len(workers) == p["workers"]
    for unit in range(p["ru"]):
        iv = sorted(covered.get(unit, []))
        assert iv and iv[0][0] == 0 and iv[-1][1] == T, (unit, iv)
        assert all(a[1] == b[0] for a, b in zip(iv, iv[1:])), f"row unit {unit}: gap or overlap {iv}"
        js = sorted(j for (u, j) in slots if u == unit)
        assert js == list(range(len(js))), f"row unit {unit}: slots {js} not dense"
        for crank in range(cs):
            rb = unit * cs + crank
            if rb >= p["num_rb"]:
                continue
            a, b = C.c_int32(), C.c_int32()
            assert lib.mcl_plan_row_block_slots(Q, V, D, sm, rb, C.byref(a), C.byref(b)) == 0
            assert a.value == rb * S * 2 and b.value == 2 * len(js)


def test_plan_fuzz(lib_built):
    """The closed-form tile plan on random shapes and SM counts (hypothesis): the properties the
    scan kernel and the merge rely on hold for every plan the host can produce."""
    from hypothesis import given, settings, strategies as st
    from multimodal_concept_learning_b200 import _lib
    lib = _lib.load()

    @settings(max_examples=120, deadline=None)
    @given(Q=st.one_of(st.integers(1, 600), st.integers(1, 40000)),
           V=st.one_of(st.integers(1, 3000), st.integers(1, 400000)),
           D=st.sampled_from([8, 64, 72, 768, 1152, 3584, 4096]),
           sm=st.one_of(st.integers(1, 12), st.sampled_from([132, 148, 160])),
           flt=st.sampled_from([0, 1, 2]))
    def run(Q, V, D, sm, flt):
        if (-(-Q // 128)) * (-(-V // 256)) > 60000:       # keep the enumeration small
            V = max(1, 60000 * 256 // (-(-Q // 128)))
        old = lib.mcl_set_option(15, flt)                  # restart charge: cold / seeded / no filter
        try:
            _check_plan(lib, _lib, Q, V, D, sm)
        finally:
            lib.mcl_set_option(15, old)

    run()


def test_plans_without_a_cold_filter_are_balanced(lib_built):
    """The seed pass (16 sample tiles) and k = 1 scans have no top-k filter to warm up: their plans
    must not trade balance for fewer segments (round 2: the seed pass of C2 took 63 us with the
    cold-filter charges -- 16 tiles on some workers, 2 on others)."""
    from multimodal_concept_learning_b200 import _lib
    lib = _lib.load()
    old = lib.mcl_set_option(15, 2)
    try:
        for Q, V, D in [(4096, 4096, 768), (1672, 262235, 1152), (32768, 4096, 1024)]:
            p = _lib.plan_scan(Q, V, D, 148)
            per = {}
            for w, unit, vt0, vt1, j, sync in _lib.plan_segments(Q, V, D, 148):
                per[w] = per.get(w, 0) + vt1 - vt0
            ideal = p["ru"] * p["num_vt"] / p["workers"]
            assert max(per.values()) <= 1.15 * ideal + 1.5, (Q, V, D, max(per.values()), ideal)
    finally:
        lib.mcl_set_option(15, old)


def test_header_is_plain_c_and_links(lib_built, tmp_path):
    """include/mcl.h is the boundary a non-Python host binds: it must compile as C99, and a C
    program linked against the .so must be able to call the host-only entry points."""
    import shutil
    import subprocess
    from multimodal_concept_learning_b200 import _lib
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    src = tmp_path / "t.c"
    src.write_text(r'''
#include <stdio.h>
#include "mcl.h"
int main(void) {
  int32_t plan[MCL_PLAN_INTS];
  if (mcl_version() != MCL_VERSION) return 1;
  if (mcl_plan_scan(8192, 152064, 3584, 148, plan) != MCL_OK) return 2;
  if (plan[0] != 64 || plan[1] != 594 || plan[2] != 56) return 3;           /* row blocks, tiles, K slices */
  if (mcl_plan_scan(0, 1, 1, 148, plan) != MCL_ERR_BAD_ARG) return 4;
  if (mcl_last_error()[0] == 0) return 5;
  if (mcl_sharded_gather_bytes(8192, 50, 8) == 0) return 6;
  printf("grid %d workers %d\\n", plan[10], plan[4]);
  return 0;
}
''')
    exe = tmp_path / "t"
    inc = os.path.join(ROOT, "include")
    libdir = os.path.dirname(_lib.LIB_PATH)
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", f"-I{inc}", str(src), "-o", str(exe),
                    f"-L{libdir}", "-l:libmcl_sm100.so", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "grid 148" in out.stdout
