"""Worker of tests/test_gpu_multi.py — launched with torch.distributed.run, one process per rank.

Row-sharded problem (each rank generates and holds only its shard); the passes exchange their partial d-vectors through the
one-shot peer-memory exchange (ciao_comm_p2p_*: deterministic, rank-ordered) — or, with CIAO_TEST_EXCHANGE=nccl, through
ncclAllReduce — and the peers' shards are attached over CUDA IPC so that the replicated SVRG++ / LFinito inner loops
TMA-prefetch remote rows.  Every rank checks its results against a single-context run of the whole problem on its own
GPU (bitwise across ranks, ≤ 1e-13 against the unsharded pass) and rank 0 prints MULTI_GPU_OK.

Two layouts:
  * one GPU per rank (gpurun --gpus 2 …): torch backend nccl; NCCL communicator + peer exchange; remote rows over NVLink;
  * CIAO_TEST_SHARE_GPU=1: all ranks on cuda:0 (the driver's 1-GPU test box) — torch backend gloo for the host-side
    handle exchange, no NCCL (it refuses two ranks on one device), the peer exchange and the row shards go through CUDA IPC
    between processes on the same device; the GPU time-slices between the processes while one waits for the other's flag."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ciao_pkg  # noqa: E402

ciao_pkg.load()
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ciaoalgorithms_jl_b200 import _lib as L  # noqa: E402
from ciaoalgorithms_jl_b200.engine import Engine  # noqa: E402
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, LFinitoSweeper, csr, interleaved_rows, shard_rows  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    share = os.environ.get("CIAO_TEST_SHARE_GPU") == "1"
    use_nccl = not share
    use_p2p = share or os.environ.get("CIAO_TEST_EXCHANGE", "p2p") == "p2p"
    if share:
        local = 0
    torch.cuda.set_device(local)
    if share:
        dist.init_process_group("gloo")
    else:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def comm_setup(eng):
        """NCCL communicator (its all-gather serves the step scalars) and/or the peer exchange (the passes' all-reduce)."""
        if use_nccl:
            obj = [Engine.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(obj, 0)
            eng.comm_init(obj[0], rank, world)
        if use_p2p:
            hs = [None] * world
            dist.all_gather_object(hs, eng.comm_p2p_handle())
            eng.comm_p2p_attach(rank, world, hs)

    def same_on_all_ranks(a):
        blobs = [None] * world
        dist.all_gather_object(blobs, np.ascontiguousarray(a).tobytes())
        return all(b == blobs[0] for b in blobs)

    N, d, seed = 6000 + 37, 1024, 0x5EED0003
    lo, hi = shard_rows(N, world, rank)
    lam = N / 100.0

    sh = Engine(local)                                  # my shard only
    sh.gen_synthetic(L.SYNTH_LASSO, N, d, seed, scale=float(N), row0=lo, n_rows=hi - lo)
    sh.set_reg(L.REG_NORML1, lam)
    comm_setup(sh)
    handles = [None] * world
    dist.all_gather_object(handles, (sh.rows_ipc_handle(), lo, hi - lo))
    sh.attach_peer_rows([h[0] for h in handles], [h[1] for h in handles], [h[2] for h in handles], rank)

    full = Engine(local)                                # the whole problem on this GPU, for comparison
    full.gen_synthetic(L.SYNTH_LASSO, N, d, seed, scale=float(N))
    full.set_reg(L.REG_NORML1, lam)

    def rel(a, b):
        return np.linalg.norm(a - b) / np.linalg.norm(b)

    x = np.random.default_rng(0).standard_normal(d) * 1e-3
    g_sh = sh.full_gradient(x, 1.0 / N)
    assert rel(g_sh, full.full_gradient(x, 1.0 / N)) < 1e-13
    assert same_on_all_ranks(g_sh)                      # rank-ordered sum: the same bits everywhere
    assert np.array_equal(g_sh, sh.full_gradient(x, 1.0 / N))   # and run to run
    assert abs(sh.objective(x)[0] - full.objective(x)[0]) < 1e-13 * full.objective(x)[0]
    assert abs(sh.max_row_sqnorm() - full.max_row_sqnorm()) == 0.0
    gamma = 1.0 / (7.0 * N * full.max_row_sqnorm())

    # SVRG++ on the sharded problem: passes all-reduced, inner epoch replicated with remote rows
    rng_a, rng_b = HostRNG(5), HostRNG(5)
    sh.svrg_init(np.zeros(d), gamma, True)
    full.svrg_init(np.zeros(d), gamma, True)
    m = N // 8
    for _ in range(3):
        sh.svrg_epoch(rng_a.rand_vec(N, m))
        full.svrg_epoch(rng_b.rand_vec(N, m))
        m *= 2
    zs, zf = sh.get_vec(L.VEC_Z_FULL), full.get_vec(L.VEC_Z_FULL)
    assert rel(zs, zf) < 1e-11, rel(zs, zf)
    # all ranks hold the same iterate bit for bit (the inner epoch is replicated, the allreduce is identical everywhere)
    assert same_on_all_ranks(zs)

    # LFinito sweeps (sequential kernel with remote rows)
    Li = np.full(N, N * full.max_row_sqnorm())
    gam = 0.999 * N / Li
    hat = 1 / np.sum(1 / gam)
    sh.lfinito_init(np.zeros(d), gam, hat)
    full.lfinito_init(np.zeros(d), gam, hat)
    dist.barrier()                                      # peers' record tails (γ_i) are written before anyone reads them
    sw = LFinitoSweeper(N, 1, 3, HostRNG(2))
    for _ in range(2):
        o = sw.next()
        sh.lfinito_outer(o, 1)
        full.lfinito_outer(o, 1)
        dist.barrier()
    assert rel(sh.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z)) < 1e-11

    # uniform shards (N divisible by the world size): the per-row step scalars c_i(z_full) of the shards are all-gathered, so
    # the replicated inner epoch runs the one-dot step with remote rows AND remote-shard scalars
    Nu = 3072 * world
    shu, fullu = Engine(local), Engine(local)
    shu.gen_synthetic(L.SYNTH_LASSO, Nu, d, seed + 2, scale=float(Nu), row0=rank * 3072, n_rows=3072)
    fullu.gen_synthetic(L.SYNTH_LASSO, Nu, d, seed + 2, scale=float(Nu))
    for eng in (shu, fullu):
        eng.set_reg(L.REG_NORML1, Nu / 100.0)
    comm_setup(shu)
    hu = [None] * world
    dist.all_gather_object(hu, shu.rows_ipc_handle())
    shu.attach_peer_rows(hu, [r * 3072 for r in range(world)], [3072] * world, rank)
    gu = 1.0 / (7.0 * Nu * fullu.max_row_sqnorm())
    rc, rd = HostRNG(11), HostRNG(11)
    shu.svrg_init(np.zeros(d), gu, False)
    fullu.svrg_init(np.zeros(d), gu, False)
    for _ in range(2):
        shu.svrg_epoch(rc.rand_vec(Nu, Nu // 2))
        fullu.svrg_epoch(rd.rand_vec(Nu, Nu // 2))
    assert rel(shu.get_vec(L.VEC_Z_FULL), fullu.get_vec(L.VEC_Z_FULL)) < 1e-11
    dist.barrier()
    shu.close()
    fullu.close()

    # replicated rows, windowed passes (bench.py --gpus G): every rank streams N/G rows, the d-vector is all-reduced and the
    # per-row step scalars c_i(z_full) of the windows are all-gathered for the replicated inner epoch
    Nw = 4096 * world
    rep, one = Engine(local), Engine(local)
    for eng in (rep, one):
        eng.gen_synthetic(L.SYNTH_LASSO, Nw, d, seed + 1, scale=float(Nw))
        eng.set_reg(L.REG_NORML1, Nw / 100.0)
    comm_setup(rep)
    rep.set_pass_window(rank * (Nw // world), Nw // world)        # collective
    g2 = 1.0 / (7.0 * Nw * one.max_row_sqnorm())
    ra, rb = HostRNG(9), HostRNG(9)
    rep.svrg_init(np.zeros(d), g2, False)
    one.svrg_init(np.zeros(d), g2, False)
    for _ in range(2):
        rep.svrg_epoch(ra.rand_vec(Nw, Nw // 4))
        one.svrg_epoch(rb.rand_vec(Nw, Nw // 4))
    assert rel(rep.get_vec(L.VEC_Z_FULL), one.get_vec(L.VEC_Z_FULL)) < 1e-11
    assert rel(rep.get_vec(L.VEC_AV), one.get_vec(L.VEC_AV)) < 1e-11
    assert same_on_all_ranks(rep.get_vec(L.VEC_Z_FULL))
    dist.barrier()
    rep.close()
    one.close()

    # the table-init passes shard with the rows (K2): every rank builds the table rows of its shard, Σ s_i is
    # all-reduced; the sequential steps need the whole table on one GPU and are refused, not silently wrong
    sh.saga_init(np.full(d, 0.01), gamma, False)
    full.saga_init(np.full(d, 0.01), gamma, False)
    assert rel(sh.get_vec(L.VEC_AV), full.get_vec(L.VEC_AV)) < 1e-12
    assert np.array_equal(sh.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z)) or rel(sh.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z)) < 1e-12
    assert np.array_equal(sh.get_table_rows(), full.get_table_rows(lo, hi - lo))
    gsh = np.linspace(0.5, 2.0, N) * (0.999 * N / (N * full.max_row_sqnorm()))      # non-uniform γ_i: each shard must take its own
    hsh = 1 / np.sum(1 / gsh)
    sh.finito_init(np.full(d, 0.01), gsh, hsh)
    full.finito_init(np.full(d, 0.01), gsh, hsh)
    assert rel(sh.get_vec(L.VEC_AV), full.get_vec(L.VEC_AV)) < 1e-12
    assert np.array_equal(sh.get_table_rows(), full.get_table_rows(lo, hi - lo))
    # static minibatches shard by row owner (Finito_basic.jl:110-118 with the batch split over the ranks, one exchange of the
    # batch's Σ in the tail kernel): batches of 512 rows, one of them straddles the shard boundary, the last one is short
    if use_p2p:
        sw_a, sw_b = BatchSweeper(N, 512, 2, HostRNG(3)), BatchSweeper(N, 512, 2, HostRNG(3))
        for _ in range(2):
            ia, pa = csr(sw_a.take(sw_a.d))
            ib, pb = csr(sw_b.take(sw_b.d))
            l0 = sh.last_timing().launches
            sh.finito_steps(ia, pa)
            # one persistent kernel for the whole epoch: the ranks' column owners exchange their sums inside it (batch.cu)
            assert sh.last_timing().launches - l0 <= 2, sh.last_timing().launches - l0
            full.finito_steps(ib, pb)
        assert rel(sh.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z)) < 1e-11, rel(sh.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z))
        assert rel(sh.get_vec(L.VEC_AV), full.get_vec(L.VEC_AV)) < 1e-11
        assert rel(sh.get_table_rows(), full.get_table_rows(lo, hi - lo)) < 1e-12
        assert same_on_all_ranks(sh.get_vec(L.VEC_Z))
        # LFinito minibatch sweeps on the shards (Finito_LFinito.jl:78-103)
        sh.lfinito_init(np.full(d, 0.01), gsh, hsh)
        full.lfinito_init(np.full(d, 0.01), gsh, hsh)
        order = np.arange(1, -(-N // 512) + 1, dtype=np.int64)
        for _ in range(2):
            sh.lfinito_outer(order, 512)
            full.lfinito_outer(order, 512)
        assert rel(sh.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z)) < 1e-11, rel(sh.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z))
        assert rel(sh.get_vec(L.VEC_AV), full.get_vec(L.VEC_AV)) < 1e-11
        assert same_on_all_ranks(sh.get_vec(L.VEC_Z))
        sh.finito_init(np.full(d, 0.01), gsh, hsh)
    if use_p2p:
        # Interleaved shards: blocks of 64 rows dealt round-robin to the ranks, so EVERY static minibatch (a multiple of
        # 64·world rows) is spread over all ranks and a sweep scales with the GPUs.  Same checks as above against the whole problem.
        B = 64
        gl = interleaved_rows(N, B, world, rank)
        shi = Engine(local)
        shi.set_row_interleave(B, rank, world)
        shi.gen_synthetic(L.SYNTH_LASSO, N, d, seed, scale=float(N), row0=rank * B, n_rows=len(gl))
        shi.set_reg(L.REG_NORML1, lam)
        comm_setup(shi)
        assert rel(shi.full_gradient(x, 1.0 / N), full.full_gradient(x, 1.0 / N)) < 1e-13
        rb = 4 * B * world
        shi.finito_init(np.full(d, 0.01), gsh, hsh)
        full.finito_init(np.full(d, 0.01), gsh, hsh)
        assert rel(shi.get_vec(L.VEC_AV), full.get_vec(L.VEC_AV)) < 1e-12
        assert np.array_equal(shi.get_table_rows(), full.get_table_rows()[gl])          # each rank took the γ_i of ITS rows
        sw_a, sw_b = BatchSweeper(N, rb, 3, HostRNG(4)), BatchSweeper(N, rb, 3, HostRNG(4))
        for _ in range(2):
            ia, pa = csr(sw_a.take(sw_a.d))
            ib, pb = csr(sw_b.take(sw_b.d))
            l0 = shi.last_timing().launches
            shi.finito_steps(ia, pa)
            assert shi.last_timing().launches - l0 <= 2
            full.finito_steps(ib, pb)
        assert rel(shi.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z)) < 1e-11, rel(shi.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z))
        assert rel(shi.get_vec(L.VEC_AV), full.get_vec(L.VEC_AV)) < 1e-11
        assert rel(shi.get_table_rows(), full.get_table_rows()[gl]) < 1e-12
        assert same_on_all_ranks(shi.get_vec(L.VEC_Z))
        shi.lfinito_init(np.full(d, 0.01), gsh, hsh)
        full.lfinito_init(np.full(d, 0.01), gsh, hsh)
        order = np.arange(1, -(-N // rb) + 1, dtype=np.int64)
        for _ in range(2):
            shi.lfinito_outer(order, rb)
            full.lfinito_outer(order, rb)
        assert rel(shi.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z)) < 1e-11, rel(shi.get_vec(L.VEC_Z), full.get_vec(L.VEC_Z))
        assert rel(shi.get_vec(L.VEC_AV), full.get_vec(L.VEC_AV)) < 1e-11
        assert same_on_all_ranks(shi.get_vec(L.VEC_Z))
        try:   # a batch that does not start at a multiple of 64·world rows has no contiguous local part
            shi.lfinito_outer(np.arange(1, -(-N // (rb + B)) + 1, dtype=np.int64), rb + B)
            raise AssertionError("misaligned minibatch on interleaved shards should fail")
        except Exception as ex:
            assert "multiple of block_rows" in str(ex), str(ex)
        shi.close()
    try:
        sh.finito_steps(np.array([1, 2], dtype=np.int64), np.array([0, 1, 2], dtype=np.int64))
        raise AssertionError("finito_steps on a shard should fail")
    except Exception as ex:
        assert "not sharded" in str(ex)
    dist.barrier()
    if rank == 0:
        print("MULTI_GPU_OK world", world, "share_gpu" if share else "one_gpu_per_rank", "p2p" if use_p2p else "nccl")
    sh.close()
    full.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
