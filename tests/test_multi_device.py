"""The peb_multi_* layer (include/pe_b200.h, pose_estimation_b200/csrc/multi.cu): several devices behind one handle,
the in-process counterpart of the one-rank-per-GPU sharding of pose_estimation_b200/multi.py (SURVEY.md 8b layer 2, 8e).

CPU part: multi.cu linked against FAKE single-device entry points (tests/host/multi_stub.cu) — sharding, hypothesis
order, concurrency of the per-device host threads, error propagation, no leaked contexts.
GPU part: the real library with several contexts on ONE device (SURVEY.md 8e "single-GPU testability"): record for
record the same bytes as peb_icp_align_batch on a single context.
"""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

from pose_estimation_b200 import multi
from pose_estimation_b200._lib import IcpParams, IcpResult
from pose_estimation_b200.testing import synth

HOST = Path(__file__).resolve().parent / "host"


# ------------------------------------------------------------------------------------------------------------
# CPU: the host layer against stubs
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def stub():
    subprocess.run(["make", "-C", str(HOST), "libpe_multistub.so"], check=True, capture_output=True)
    L = C.CDLL(str(HOST / "libpe_multistub.so"))
    vp, sz, i = C.c_void_p, C.c_size_t, C.c_int
    L.peb_multi_create.argtypes = [i, C.POINTER(i), C.POINTER(vp)]
    L.peb_multi_destroy.argtypes = [vp]
    L.peb_multi_destroy.restype = None
    L.peb_multi_last_error.argtypes = [vp]
    L.peb_multi_last_error.restype = C.c_char_p
    L.peb_multi_size.argtypes = [vp]
    L.peb_multi_ctx.argtypes = [vp, i]
    L.peb_multi_ctx.restype = vp
    L.peb_multi_set_int.argtypes = [vp, C.c_char_p, i]
    L.peb_multi_shard_range.argtypes = [sz, i, i, C.POINTER(sz), C.POINTER(sz)]
    L.peb_multi_shard_range.restype = None
    L.peb_multi_target_set.argtypes = [vp, vp, sz, sz, vp, sz]
    L.peb_multi_source_set.argtypes = [vp, vp, sz, sz]
    L.peb_multi_icp_align_batch.argtypes = [vp, vp, sz, C.POINTER(IcpParams), vp]
    L.peb_multi_launch_count.argtypes = [vp]
    L.peb_multi_launch_count.restype = C.c_uint64
    L.stub_ctx_batch_streams.argtypes = [vp]
    return L


def _create(L, devices):
    L.stub_reset_ordinals()
    h = C.c_void_p()
    arr = (C.c_int * len(devices))(*devices)
    rc = L.peb_multi_create(len(devices), arr, C.byref(h))
    return rc, h


def _align(L, h, H, params=None):
    g = np.zeros((max(H, 1), 16), np.float32)
    g[:, 0] = 100.0 + np.arange(max(H, 1))
    res = (IcpResult * max(H, 1))()
    prm = params or IcpParams(max_iterations=7)
    rc = L.peb_multi_icp_align_batch(h, g.ctypes.data, H, C.byref(prm), res)
    return rc, res


@pytest.mark.parametrize("n,world", [(0, 1), (1, 4), (13, 2), (13, 3), (96, 8), (1024, 8), (1000, 7), (5, 64)])
def test_c_shard_rule_is_the_python_shard_rule(stub, n, world):
    lo, hi = C.c_size_t(), C.c_size_t()
    for r in range(world):
        stub.peb_multi_shard_range(n, world, r, C.byref(lo), C.byref(hi))
        assert (lo.value, hi.value) == multi.shard_range(n, world, r)
    stub.peb_multi_shard_range(n, world, world, C.byref(lo), C.byref(hi))  # out of range: empty
    assert (lo.value, hi.value) == (0, 0)


@pytest.mark.parametrize("devices,H", [([0], 5), ([0, 1, 2], 13), ([0, 1, 2, 3], 2), ([3, 3], 9), (list(range(8)), 96)])
def test_every_hypothesis_is_refined_once_in_order_by_the_device_of_its_block(stub, devices, H):
    rc, h = _create(stub, devices)
    assert rc == 0 and stub.peb_multi_size(h) == len(devices) and stub.stub_live_contexts() == len(devices)
    pts = np.zeros((11, 4), np.float32)
    assert stub.peb_multi_target_set(h, pts.ctypes.data, 11, 16, None, 0) == 0
    assert stub.peb_multi_source_set(h, pts.ctypes.data, 5, 16) == 0
    stub.stub_max_concurrent()
    rc, res = _align(stub, h, H)
    assert rc == 0
    busy = sum(1 for r in range(len(devices)) if multi.shard_range(H, len(devices), r)[0] < multi.shard_range(H, len(devices), r)[1])
    assert stub.stub_max_concurrent() == busy  # the shards run at the same time, one host thread per device
    for r in range(len(devices)):
        lo, hi = multi.shard_range(H, len(devices), r)
        for k in range(lo, hi):
            assert res[k].state == r and res[k].n_correspondences == devices[r]  # which context, which device
            assert res[k].T[0] == 100.0 + k and res[k].iterations == 7          # its own guess, the caller's parameters
            assert res[k].fitness == 5 + 1e-3 * 11                              # every replica saw the source and the target
    assert stub.peb_multi_launch_count(h) == 2 * len(devices) + H
    stub.peb_multi_destroy(h)
    assert stub.stub_live_contexts() == 0


def test_scene_and_model_cross_pcie_once_and_the_workers_are_persistent(stub):
    """peb_multi_target_set / _source_set: ONE host-to-device staging (context 0) + a device-to-device clone per extra
    context; the per-device host threads are created with the handle and reused by every call."""
    stub.stub_uploads.restype = C.c_int
    stub.stub_clones.restype = C.c_int
    stub.stub_ctx_thread.argtypes = [C.c_void_p]
    stub.stub_ctx_ready.argtypes = [C.c_void_p]
    rc, h = _create(stub, [0, 1, 2, 3])
    assert rc == 0
    stub.stub_uploads(), stub.stub_clones()
    pts = np.zeros((11, 4), np.float32)
    assert stub.peb_multi_target_set(h, pts.ctypes.data, 11, 16, None, 0) == 0
    assert (stub.stub_uploads(), stub.stub_clones()) == (1, 3)
    assert stub.peb_multi_source_set(h, pts.ctypes.data, 5, 16) == 0
    assert (stub.stub_uploads(), stub.stub_clones()) == (1, 3)
    assert [stub.stub_ctx_ready(stub.peb_multi_ctx(h, i)) for i in range(4)] == [3, 3, 3, 3]
    threads = []
    for _ in range(3):
        rc, _res = _align(stub, h, 8)
        assert rc == 0
        threads.append([stub.stub_ctx_thread(stub.peb_multi_ctx(h, i)) for i in range(4)])
    assert threads[0] == threads[1] == threads[2]  # the same host thread drives a device in every call
    assert len(set(threads[0])) == 4               # and every device has its own
    stub.peb_multi_destroy(h)
    assert stub.stub_live_contexts() == 0


def test_empty_batch_and_bad_arguments(stub):
    rc, h = _create(stub, [0, 1])
    assert rc == 0
    res = (IcpResult * 1)()
    prm = IcpParams()
    assert stub.peb_multi_icp_align_batch(h, None, 0, C.byref(prm), res) == 0      # nothing to do
    assert stub.peb_multi_icp_align_batch(h, None, 3, C.byref(prm), res) == -1     # PEB_E_INVALID_ARG
    assert b"null guesses" in stub.peb_multi_last_error(h)
    assert stub.peb_multi_icp_align_batch(None, None, 3, C.byref(prm), res) == -1
    assert stub.peb_multi_set_int(h, b"batch_streams", 3) == 0
    assert [stub.stub_ctx_batch_streams(stub.peb_multi_ctx(h, i)) for i in range(2)] == [3, 3]
    assert stub.peb_multi_set_int(h, b"no_such_knob", 1) == -1 and b"no_such_knob" in stub.peb_multi_last_error(h)
    assert stub.peb_multi_ctx(h, 2) is None and stub.peb_multi_ctx(h, -1) is None
    stub.peb_multi_destroy(h)
    stub.peb_multi_destroy(None)
    assert stub.stub_live_contexts() == 0


def test_a_failing_device_is_reported_and_the_other_shards_still_finish(stub):
    rc, h = _create(stub, [0, 1, 2])
    assert rc == 0
    stub.stub_fail_ordinal(1)
    try:
        rc, res = _align(stub, h, 9)
    finally:
        stub.stub_fail_ordinal(-1)
    assert rc == -4  # PEB_E_CUDA, the failing context's own status
    msg = stub.peb_multi_last_error(h).decode()
    assert "context 1" in msg and "device 1" in msg and "injected failure" in msg
    assert [res[k].state for k in (0, 1, 2, 6, 7, 8)] == [0, 0, 0, 2, 2, 2]
    rc, _ = _align(stub, h, 9)  # the handle stays usable
    assert rc == 0
    stub.peb_multi_destroy(h)
    assert stub.stub_live_contexts() == 0


def test_create_fails_as_a_whole_and_leaks_nothing(stub):
    rc, h = _create(stub, [0, 1, 99])
    assert rc == -1 and not h.value and stub.stub_live_contexts() == 0
    msg = stub.peb_multi_last_error(None).decode()
    assert "context 2" in msg and "device 99" in msg
    h = C.c_void_p()
    assert stub.peb_multi_create(0, None, C.byref(h)) == -1 and stub.peb_multi_create(65, None, C.byref(h)) == -1
    assert stub.peb_multi_create(2, None, None) == -1
    rc = stub.peb_multi_create(3, None, C.byref(h))  # NULL device list = 0 .. ndev-1
    assert rc == 0 and stub.peb_multi_size(h) == 3
    stub.peb_multi_destroy(h)
    assert stub.stub_live_contexts() == 0


def test_real_library_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from pose_estimation_b200 import pcl

    with pytest.raises(pcl.PebError) as e:
        pcl.MultiContext([0, 1])
    assert "no CPU fallback" in str(e.value) and "context 0" in str(e.value)


# ------------------------------------------------------------------------------------------------------------
# GPU: several contexts on one device against a single context
# ------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def pcl():
    from pose_estimation_b200 import pcl as m

    return m


@pytest.fixture(scope="module")
def problem(oracle):
    return synth.make_c2(scale=0.25, downsample=lambda p, leaf: oracle.voxel_grid(p, leaf)[0])


def _params(icp, **kw):
    from oracle import default_params

    prm = default_params(**kw)
    for name, _ in prm._fields_:
        setattr(icp.params, name, getattr(prm, name))


@pytest.mark.gpu
@pytest.mark.parametrize("devices,H", [([0, 0], 13), ([0, 0, 0], 2), ([0, 0], 48)])
def test_sharded_batch_is_bit_identical_to_the_single_context_batch(pcl, problem, devices, H):
    """13 -> ragged blocks of 7 + 6; 2 hypotheses on 3 contexts -> one context idle; 48 -> two blocks of 24, large
    enough for the per-hypothesis launch dependencies, running concurrently on the one device."""
    p = problem
    rng = np.random.default_rng(123)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006) for _ in range(H)])
    out = []
    for ctx in (pcl.Context(0), pcl.MultiContext(devices)):
        icp = pcl.IterativeClosestPoint(ctx)
        icp.setInputSource(p.source)
        icp.setInputTarget(p.target)
        _params(icp, max_iterations=15, max_corr_dist=0.02)
        out.append([bytes(r) for r in icp.alignBatch(guesses)])
        if getattr(ctx, "is_multi", False):
            assert ctx.size == len(devices) and ctx.launch_count > 0
            icp.align(guesses[0], want_output=False)  # not sharded: runs on the first context
            assert bytes(icp.result) == out[0][0]
        ctx.close()
    assert out[0] == out[1]


@pytest.mark.gpu
def test_sharded_point_to_plane_batch_and_errors(pcl, problem):
    p = problem
    rng = np.random.default_rng(9)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 4.0, 0.004) for _ in range(10)])
    single, many = pcl.Context(0), pcl.MultiContext([0, 0])
    nrm = pcl.NormalEstimation(single)
    nrm.setInputCloud(p.target)
    nrm.setKSearch(12)
    normals = nrm.compute()
    out = []
    for ctx in (single, many):
        icp = pcl.IterativeClosestPointWithNormals(ctx)
        icp.setInputSource(p.source)
        icp.setInputTarget(p.target, normals)
        _params(icp, max_iterations=10, max_corr_dist=0.02)
        icp.params.estimator = 1
        out.append([bytes(r) for r in icp.alignBatch(guesses)])
    assert out[0] == out[1]
    fresh = pcl.MultiContext([0, 0])
    icp = pcl.IterativeClosestPoint(fresh)
    with pytest.raises(pcl.PebError) as e:  # no target, no source: the device's own status comes through
        icp.alignBatch(guesses)
    assert e.value.code in (-2, -3) and "context 0" in str(e.value)
    with pytest.raises(pcl.PebError):
        pcl.MultiContext([0, 4096])
    for c in (single, many, fresh):
        c.close()


@pytest.mark.gpu
def test_two_devices_when_the_box_has_them(pcl, problem):
    try:
        many = pcl.MultiContext([0, 1])
    except pcl.PebError:
        pytest.skip("one GPU only")
    p = problem
    rng = np.random.default_rng(31)
    guesses = np.stack([synth.perturb_pose(p.gt_pose, rng, 5.0, 0.006) for _ in range(40)])

    def run(ctx, g):
        icp = pcl.IterativeClosestPoint(ctx)
        icp.setInputSource(p.source)
        icp.setInputTarget(p.target)
        _params(icp, max_iterations=15, max_corr_dist=0.02)
        return [bytes(r) for r in icp.alignBatch(g)]

    sharded = run(many, guesses)
    many.close()
    # the contract (pe_b200.h): every block exactly as one context refines that block — here one context per device,
    # so the replica built from the device-to-device clone on device 1 is checked against a grid built from the host buffer
    blocks = []
    for r, dev in enumerate((0, 1)):
        lo, hi = multi.shard_range(len(guesses), 2, r)
        ctx = pcl.Context(dev)
        blocks += run(ctx, guesses[lo:hi])
        ctx.close()
    assert sharded == blocks
    # and against one context refining all 40 (same block partition at this size: identical as well)
    ctx = pcl.Context(0)
    assert sharded == run(ctx, guesses)
    ctx.close()
