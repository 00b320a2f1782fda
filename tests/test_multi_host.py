"""Host-side multi-GPU logic on the CPU: shard bookkeeping and the gather layout with a
world_size-2 (and 3) gloo group; the C ABI library loads and exports what the header declares."""
import ctypes as C
import os
import re
import socket
from pathlib import Path

import numpy as np
import pytest

from pose_estimation_b200 import multi

ROOT = Path(__file__).resolve().parents[1]


def test_record_layout_matches_the_c_struct():
    from oracle import IcpResult

    assert multi.RECORD_BYTES == C.sizeof(IcpResult) == 96
    for name, (dtype, offset) in multi.RESULT_DTYPE.fields.items():
        cname = {"last_mse": "last_mse"}.get(name, name)
        assert getattr(IcpResult, cname).offset == offset


@pytest.mark.parametrize("n,world", [(1024, 1), (1024, 2), (1024, 8), (1000, 8), (5, 8), (0, 2)])
def test_shard_ranges_partition_the_hypotheses(n, world):
    ranges = [multi.shard_range(n, world, r) for r in range(world)]
    covered = [i for lo, hi in ranges for i in range(lo, hi)]
    assert covered == list(range(n))
    assert all(hi - lo <= multi.shard_size(n, world) for lo, hi in ranges)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_items, q):
    import torch
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = multi.shard_range(n_items, world, rank)
        recs = np.zeros(hi - lo, multi.RESULT_DTYPE)
        for k, h in enumerate(range(lo, hi)):
            recs[k]["T"] = np.arange(16, dtype=np.float32) + h
            recs[k]["fitness"] = 1.0 / (1 + (h * 7919) % 101)
            recs[k]["iterations"] = 30
            recs[k]["converged"] = 1 if h % 5 else 0
            recs[k]["n_correspondences"] = h
        local = torch.from_numpy(np.frombuffer(recs.tobytes(), dtype=np.uint8).copy())
        gathered = multi.gather_results(local, n_items, world, rank)
        out = multi.unpack_results(gathered.numpy().tobytes(), n_items, world)
        q.put((rank, out["n_correspondences"].tolist(), float(out["fitness"].sum()), multi.best_hypothesis(out)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 37), (3, 10)])
def test_gloo_allgather_of_result_records(world, n_items):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect_fit = [1.0 / (1 + (h * 7919) % 101) for h in range(n_items)]
    expect_best = int(np.argmin([f if h % 5 else np.inf for h, f in enumerate(expect_fit)]))
    for rank, ncorr, fsum, best in got:
        assert ncorr == list(range(n_items))          # every rank holds all records, in hypothesis order
        assert abs(fsum - sum(expect_fit)) < 1e-12
        assert best == expect_best                    # so the best pose is chosen locally and identically


def test_library_exports_every_symbol_the_header_declares():
    header = (ROOT / "include" / "pe_b200.h").read_text()
    declared = set(re.findall(r"PEB_API\s+[\w\s\*]+?\b(peb_\w+)\s*\(", header))
    assert len(declared) >= 25
    from pose_estimation_b200 import _lib

    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    lib = C.CDLL(str(_lib.LIB_PATH))
    for name in declared:
        assert hasattr(lib, name), name
    assert b"sm_100a" in _lib.lib.peb_version()


def test_no_gpu_means_a_loud_failure_not_a_fallback():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from pose_estimation_b200 import pcl

    with pytest.raises(pcl.PebError) as e:
        pcl.Context(0)
    assert "no CPU fallback" in str(e.value)


def test_product_never_touches_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import, link or call oracle/."""
    pat = re.compile(r"import\s+oracle|from\s+oracle|libpcl_oracle|\borc_\w+\s*\(|pcl_oracle\.cpp\"|dlopen")
    for base in (ROOT / "pose_estimation_b200", ROOT / "include"):
        for path in base.rglob("*"):
            if path.suffix in (".py", ".cu", ".cuh", ".h", ".hpp", ".cpp") or path.name == "Makefile":
                assert not pat.search(path.read_text()), path


def test_header_is_plain_c99_and_the_library_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: the header must compile as C (pedantic C99), the library must link without a
    C++ driver, the POD layouts are fixed, and with no GPU the entry points refuse instead of computing on the CPU."""
    import subprocess

    exe = tmp_path / "abi_check"
    lib_dir = ROOT / "pose_estimation_b200"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"),
                        str(ROOT / "tests" / "c" / "abi_check.c"), "-o", str(exe), "-L", str(lib_dir), "-lpe_b200", "-lm",
                        f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    import torch

    if torch.cuda.is_available():
        assert r.returncode == 0 and "ctx ok" in r.stdout, r.stdout
    else:
        assert r.returncode == 10, r.stdout
        assert "no CPU fallback" in r.stdout and "multi_create" in r.stdout
