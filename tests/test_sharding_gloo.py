"""N>1 host logic on CPU: two gloo ranks shard the golden reads contiguously, classify their shard (kernel logic on the CPU
warp emulator -- there is no GPU here), all-gather the taxon arrays and must reproduce the reference output in input order."""
import os, subprocess, sys, textwrap
import numpy as np
import pytest
from conftest import ROOT

WORKER = textwrap.dedent('''
    import os, sys, ctypes as C
    import numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
    from conftest import Golden
    from helpers import make_params
    from test_kernel_logic_emulated import KjParams, emu_classify
    from kaiju_b200.sharding import shard_bounds, slice_packed, all_gather_taxa, all_reduce_counts
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%(port)d", rank=int(sys.argv[1]), world_size=2)
    rank, world = dist.get_rank(), dist.get_world_size()
    g = Golden(); names, s1, o1, s2, o2 = g.reads("pe150"); n = len(names)
    lo, hi = shard_bounds(n, rank, world)
    a1, b1 = slice_packed(s1, o1, lo, hi); a2, b2 = slice_packed(s2, o2, lo, hi)
    E = C.CDLL(os.path.join(%(root)r, "tests", "emu", "libkjemu.so"))
    E.kjemu_create.restype = C.c_void_p; E.kjemu_create.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(KjParams)]
    E.kjemu_destroy.argtypes = [C.c_void_p]; E.kjemu_classify.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_uint64, C.c_void_p, C.c_void_p, C.c_int]
    tax, best = emu_classify(E, g.fmi, g.nodes, make_params("mem"), np.ascontiguousarray(a1), b1, np.ascontiguousarray(a2), b2)
    full = all_gather_taxa(torch.from_numpy(tax.view(np.int64)), n, rank, world, dist)
    etax, _, _ = g.expected("mem_default", "pe150")
    ok = np.array_equal(full.numpy().view(np.uint64), etax)
    # per-taxon summary: every rank histograms its shard over a common id table, one all-reduce gives the job-wide counts
    table = np.unique(etax); local = np.zeros(len(table), dtype=np.int64)
    u, c = np.unique(tax, return_counts=True); local[np.searchsorted(table, u)] = c
    total = all_reduce_counts(torch.from_numpy(local), dist).numpy()
    eu, ec = np.unique(etax, return_counts=True)
    ok = ok and np.array_equal(total, ec) and int(total.sum()) == n
    print("RANK", rank, "shard", lo, hi, "OK" if ok else "MISMATCH", flush=True)
    dist.barrier(); dist.destroy_process_group()
    sys.exit(0 if ok else 1)
''')


def test_shard_bounds_cover_and_balance():
    from kaiju_b200.sharding import shard_bounds
    for n in (0, 1, 7, 1520, 10_000_000):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_two_rank_gloo_sharded_classification(built, tmp_path):
    port = 29500 + (os.getpid() % 1000)
    script = tmp_path / "worker.py"; script.write_text(WORKER % {"root": ROOT, "port": port})
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=600)[0].decode() for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert all("OK" in o for o in outs), outs
