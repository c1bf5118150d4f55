"""Trial-sharded mode over NCCL (the real exchange inside libmcp.so -- slices by grouped send / recv, or the all-gather of
whole blocks): two processes, one GPU each.
Needs a box with >= 2 GPUs (`gpurun --gpus 2`); on a 1-GPU box the exchange kernels are still covered by
tests/test_gpu_trial_sharded.py, which emulates the ranks."""
import numpy as np
import pytest

import riskcases
from mcp_b200 import native as N
from mcp_b200 import sharding
from oracle import cref

pytestmark = pytest.mark.gpu
CID = N.CODEC_FASTPFOR256_VB | N.FLAG_DELTA


def _gpu_count() -> int:
    import ctypes
    try:
        rt = ctypes.CDLL("libcudart.so")
    except OSError:
        try:
            rt = ctypes.CDLL("/usr/local/cuda/lib64/libcudart.so")
        except OSError:
            return 0
    n = ctypes.c_int(0)
    return n.value if rt.cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0


def _worker(rank, world, uid_q, out_q, case):
    import mcp_b200
    E, S, alpha, exchange = case
    try:
        ctx = mcp_b200.Context(rank)
        ctx.set_option("ts_exchange", exchange)
        if rank == 0:
            uid = ctx.comm_unique_id()
            for _ in range(world - 1):
                uid_q.put(uid)
        else:
            uid = uid_q.get(timeout=120)
        n0, n1 = sharding.trial_shard(S.shape[0], rank, world)
        r = sharding.run_trial_sharded_gpu(ctx, E, S[n0:n1], alpha, S.shape[0], n0, rank, world, CID, N.FRAME_APP, uid)
        ctx.close()
        out_q.put((rank, r))
    except Exception as e:  # surface the failure instead of hanging the parent
        out_q.put((rank, repr(e)))


def _run(world, case):
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    uid_q, out_q = mpc.Queue(), mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, world, uid_q, out_q, case)) for r in range(world)]
    for p in procs:
        p.start()
    parts = [out_q.get(timeout=300) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, r in parts:
        assert isinstance(r, dict), f"rank {rank}: {r}"
    return sharding.assemble([r for _, r in parts], case[0].shape[0]), [r["fallback_rows"] for _, r in parts]


@pytest.mark.skipif(_gpu_count() < 2, reason="needs 2 GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("exchange", [1, 0], ids=["slices", "allgather"])
def test_nccl_allgather_trial_sharded_two_gpus(exchange):
    """ts_exchange = 1 (default): a rank receives only its portfolio slice of the other ranks' blocks (131 rows: slices of
    66 and 65); 0: the all-gather of whole blocks.  Same results, and the oracle's."""
    E, S = riskcases.synthetic(131, 60000, 32, 71)
    got, fb = _run(2, (E, S, 0.99, exchange))
    o = cref.risk(E, S, 0.99)
    assert fb == [0, 0]
    assert np.array_equal(got["var_trial"], o["var_trial"]) and np.array_equal(got["var_loss"], o["var_loss"])
    assert np.array_equal(got["breach"], o["breach"])
    for p in (0, 64, 65, 66, 130):
        assert bytes(got["enc"][got["enc_off"][p]: got["enc_off"][p + 1]]) == cref.get_compressed_instruments(CID, o["breach"][p])
    np.testing.assert_allclose(got["variance"], o["variance"], rtol=1e-9, atol=0)
    np.testing.assert_allclose(got["cvar"], o["cvar"], rtol=1e-9, atol=0)
    # adversarial order: the whole tail lives on one rank -> second all-gather round, still exact
    Nn, F, P = 40000, 32, 16
    S2 = np.zeros((Nn, F))
    S2[:, 0] = -np.linspace(0.0, 1.0, Nn)
    S2[:, 1] = 1e-3 * np.cos(np.arange(Nn))
    E2 = np.abs(riskcases.synthetic(P, 1, F, 3)[0]) + 1.0
    got, fb = _run(2, (E2, S2, 0.99, exchange))
    o = cref.risk(E2, S2, 0.99)
    assert fb == [P, P]
    assert np.array_equal(got["breach"], o["breach"]) and np.array_equal(got["var_trial"], o["var_trial"])
