"""Host-side logic: index selection quirks of the reference, operator recognition, row sharding,
and the N>1 exchange step emulated with gloo on CPU (world_size 2)."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)       # spawned gloo workers re-import this module without conftest
import ciao_pkg  # noqa: E402

ciao_pkg.load()

from ciaoalgorithms_jl_b200 import operators as ops  # noqa: E402
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, LFinitoSweeper, csr, shard_rows, static_batches  # noqa: E402


def test_cyclic_sweep_starts_at_batch_two():                      # Finito_basic.jl:38,99
    sw = BatchSweeper(6, 1, 2, HostRNG(0))
    assert [int(b[0]) for b in sw.take(7)] == [2, 3, 4, 5, 6, 1, 2]


def test_shuffled_sweep_first_pass_is_natural_order():            # Finito_basic.jl:39-40,101-107
    sw = BatchSweeper(5, 1, 3, HostRNG(0))
    first = [int(b[0]) for b in sw.take(5)]
    second = sorted(int(b[0]) for b in sw.take(5))
    assert first == [1, 2, 3, 4, 5] and second == [1, 2, 3, 4, 5]


def test_static_batches_with_remainder():                         # Finito_basic.jl:52-57
    b = static_batches(7, 3)
    assert [list(x) for x in b] == [[1, 2, 3], [4, 5, 6], [7]]
    sw = BatchSweeper(7, 3, 2, HostRNG(0))
    assert sw.d == 3 and list(sw.next()) == [4, 5, 6] and list(sw.next()) == [7] and list(sw.next()) == [1, 2, 3]


def test_random_minibatch_without_replacement():                  # Finito_basic.jl:97
    sw = BatchSweeper(10, 4, 1, HostRNG(3))
    for b in sw.take(20):
        assert len(set(b.tolist())) == 4 and b.min() >= 1 and b.max() <= 10


def test_lfinito_sweeper():                                       # Finito_LFinito.jl:89-91
    assert list(LFinitoSweeper(5, 2, 2, HostRNG(0)).next()) == [1, 2, 3]
    assert list(LFinitoSweeper(5, 2, 1, HostRNG(0)).next()) == [1, 2, 3]          # sweeping 1 is silently cyclic
    assert sorted(LFinitoSweeper(5, 2, 3, HostRNG(0)).next()) == [1, 2, 3]


def test_csr():
    idx, ptr = csr([np.array([3, 1]), np.array([2]), np.array([], dtype=np.int64)])
    assert idx.tolist() == [3, 1, 2] and ptr.tolist() == [0, 2, 3, 3]


def test_shard_rows_partition():
    for N, G in [(10, 3), (1 << 22, 8), (7, 8)]:
        bounds = [shard_rows(N, G, r) for r in range(G)]
        assert bounds[0][0] == 0 and bounds[-1][1] == N
        assert all(bounds[r][1] == bounds[r + 1][0] for r in range(G - 1))


def test_operator_recognition():
    F = [ops.LeastSquares(np.array([[1.0, 2.0, 3.0]]), np.array([0.5]), 6.0) for _ in range(4)]
    kind, loss, A, b, s = ops.pack_F(F, 4)
    assert kind == "rows" and loss == 0 and A.shape == (4, 3) and np.all(b == 0.5) and np.all(s == 6.0)
    F = [ops.Precompose(ops.LogisticLoss(np.array([-1.0]), 1.0), np.ones((1, 5)), 1.0) for _ in range(3)]
    kind, loss, A, y, mu = ops.pack_F(F, 3)
    assert loss == 1 and np.all(y == -1.0) and A.shape == (3, 5)
    box = ops.IndBox(-2.0, 2.0)
    F = [ops.Sum(ops.Quadratic(np.diag([1.0, 2.0]), np.ones(2)), ops.SqrDistL2(box, 30.0)) for _ in range(3)]
    kind, Qd, ql, bx, eta = ops.pack_F(F, 3)
    assert kind == "blocks" and Qd.shape == (3, 2) and bx == (-2.0, 2.0) and eta == 30.0
    with pytest.raises(ops.UnsupportedOperator):
        ops.pack_F([ops.Sum(ops.Quadratic(np.array([[1.0, 0.5], [0.5, 1.0]]), np.ones(2)), ops.SqrDistL2(box, 1.0))], 1)
    with pytest.raises(ops.UnsupportedOperator):
        ops.pack_F([object()], 1)
    assert ops.reg_params(ops.NormL1(0.3)) == (1, 0.3) and ops.reg_params(ops.Zero()) == (0,)


# ---------------------------------------------------------------------------------------------
def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    """The sharded pass of SURVEY.md §8e on CPU: each rank evaluates its contiguous row shard with the
    oracle, one all_reduce(sum) of the d-vector, result must equal the unsharded pass."""
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    import ciao_pkg
    ciao_pkg.load()
    from ciaoalgorithms_jl_b200.sampling import shard_rows as sr
    from oracle import oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, d, seed = 301, 48, 99
    lo, hi = sr(N, world, rank)
    A, b = orc.gen_rows(orc.SYN_LASSO, d, seed, lo, hi - lo)      # each shard generates its own rows
    x = np.linspace(-1, 1, d)
    part = orc.Problem(orc.LOSS_LS, A, b, np.full(hi - lo, float(N))).full_gradient(x, 1.0)
    t = torch.from_numpy(part.copy())
    dist.all_reduce(t)
    fs = torch.tensor([orc.Problem(orc.LOSS_LS, A, b, np.full(hi - lo, float(N))).objective(x)[0] * (hi - lo)], dtype=torch.float64)
    dist.all_reduce(fs)
    if rank == 0:
        Af, bf = orc.gen_rows(orc.SYN_LASSO, d, seed, 0, N)
        full = orc.Problem(orc.LOSS_LS, Af, bf, np.full(N, float(N)))
        q.put((np.abs(t.numpy() / N - full.full_gradient(x, 1.0 / N)).max() / np.abs(full.full_gradient(x, 1.0 / N)).max(),
               abs(fs.item() / N - full.objective(x)[0]) / full.objective(x)[0]))
    dist.destroy_process_group()


def test_sharded_pass_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err_g, err_f = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err_g < 1e-13 and err_f < 1e-13


def test_adaptive_sweeper_follows_the_reference():                # Finito_adaptive.jl:53-55, 107-119
    from ciaoalgorithms_jl_b200.sampling import AdaptiveSweeper, HostRNG
    cyc = AdaptiveSweeper(5, 2, HostRNG(0)).take(7)
    assert cyc.tolist() == [1, 2, 3, 4, 5, 1, 2]                  # idxr starts at 0: the first index is 1 (the basic variant starts at 2)
    shf = AdaptiveSweeper(5, 3, HostRNG(0)).take(10)
    assert shf[:5].tolist() == [1, 2, 3, 4, 5] and sorted(shf[5:].tolist()) == [1, 2, 3, 4, 5]   # natural first pass, then randperm
    rnd = AdaptiveSweeper(5, 1, HostRNG(0)).take(50)
    assert rnd.min() >= 1 and rnd.max() <= 5


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU restatement on the host cores) prints one JSON line with the contract's keys."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-rows-log2", "11"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["unit"] == "epochs/s" and line["value"] > 0 and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert "workload" in line["config"] and "model" not in line["config"]


def test_bench_reference_arm_runs_one_solve_per_gpu_on_as_many_cores():
    """At N > 1 the GPU arm runs N independent solves; the reference arm runs the same N solves concurrently, one per core."""
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0",
                          "--cpu-rows-log2", "11"], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["scaling"] == "weak"
    assert line["cpu_baseline"]["cores"] == 2 and "2 independent solves" in line["config"]["workload"]
    # a rank other than 0 prints nothing and exits 0 (torchrun launches the arm on every rank)
    res = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=60, env=dict(os.environ, RANK="1"))
    assert res.returncode == 0 and res.stdout.strip() == ""


# ---- Julia's default RNG, restated (csrc/jlrng.cu, sampling.JuliaRNG) -----------------------------------------------
def test_julia_rng_known_answer_from_the_julia_docs():
    """`Random.seed!(1234); rand(2)` in Julia's documentation of Random.seed! / Xoshiro (Julia ≥ 1.7):
    0.32597672886359486, 0.5490511363155669 — pins seeding (SHA-256 of the seed words), xoshiro256++ and the Float64 conversion."""
    from ciaoalgorithms_jl_b200.sampling import JuliaRNG
    x = JuliaRNG(1234).rand_float(2)
    assert x[0] == 0.32597672886359486 and x[1] == 0.5490511363155669


def test_julia_rng_samplers_are_well_formed_and_follow_the_stream():
    from ciaoalgorithms_jl_b200.sampling import JuliaRNG
    M64 = (1 << 64) - 1
    # rand(1:N): Lemire's nearly-divisionless method on the UInt64 stream, re-derived here in Python integers
    a, b = JuliaRNG(7), JuliaRNG(7)
    N = 1000003
    got = a.rand_vec(N, 2000)
    raw = iter(int(v) for v in b.next_u64(4000))
    want = []
    for _ in range(2000):
        m = next(raw) * N
        if (m & M64) < N:
            t = ((1 << 64) - N) % N
            while (m & M64) < t:
                m = next(raw) * N
        want.append((m >> 64) + 1)
    assert got.tolist() == want and got.min() >= 1 and got.max() <= N
    # randperm / sample without replacement: permutations, distinct, in range, reproducible, state carried over between calls
    r = JuliaRNG(3)
    p = r.randperm(1000)
    assert sorted(p.tolist()) == list(range(1, 1001))
    assert not np.array_equal(p, r.randperm(1000))
    assert np.array_equal(JuliaRNG(3).randperm(1000), p)
    for N, k in ((50, 1), (50, 2), (50, 7), (100000, 5), (10, 10)):      # k = 1 | samplepair | Fisher–Yates | self-avoiding | all
        s = r.sample_norep(N, k)
        assert len(set(s.tolist())) == k and s.min() >= 1 and s.max() <= N
    # the samplers plug into the reference's index-selection logic
    sw = BatchSweeper(10, 3, 3, JuliaRNG(0))
    seen = [b for _ in range(2 * sw.d) for b in sw.next().tolist()]
    assert sorted(seen[:10]) == list(range(1, 11)) and sorted(seen[10:]) == list(range(1, 11))


def test_interleaved_shards_give_every_aligned_batch_a_contiguous_local_window():
    """Interleaved row shards (ciao_set_row_interleave): rank k holds the blocks of B rows number k, k + G, …  A static batch that
    starts at a multiple of B·G rows must map to ONE contiguous range of local rows on every rank — [lo/G, lo/G + count) with the
    count formula of csrc/common.cuh il_count — and the ranks' parts must tile the batch."""
    from ciaoalgorithms_jl_b200.sampling import interleaved_rows

    def il_count(n, B, rank, world):
        sb = B * world
        rem = n % sb - rank * B
        return (n // sb) * B + min(max(rem, 0), B)

    for N, B, G in ((6037, 64, 2), (6037, 64, 3), (1 << 14, 256, 8), (1000, 7, 4)):
        owned = [interleaved_rows(N, B, G, k) for k in range(G)]
        assert sorted(np.concatenate(owned).tolist()) == list(range(N))
        for k in range(G):
            assert len(owned[k]) == il_count(N, B, k, G)
            assert owned[k][0] == k * B or len(owned[k]) == 0
        r = 4 * B * G
        for lo in range(0, N, r):
            n = min(r, N - lo)
            total = 0
            for k in range(G):
                loc = np.nonzero((owned[k] >= lo) & (owned[k] < lo + n))[0]
                cnt = il_count(n, B, k, G)
                assert len(loc) == cnt
                if cnt:
                    assert loc[0] == lo // G and loc[-1] == lo // G + cnt - 1      # contiguous local rows
                total += cnt
            assert total == n


def _minibatch_worker(rank, world, port, q):
    """One rank of a Finito minibatch epoch on interleaved row shards, emulated with the oracle's component gradients and a gloo
    all-reduce per batch: every rank updates the table rows it owns and contributes its part of the batch's Σ (Finito_basic.jl:110-118)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import ciao_pkg
    ciao_pkg.load()
    from ciaoalgorithms_jl_b200.sampling import interleaved_rows, static_batches
    from oracle import oracle as orc
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    N, d, seed, B = 192, 24, 7, 8
    A, b = orc.gen_rows(orc.SYN_LASSO, d, seed, 0, N)
    prob = orc.Problem(orc.LOSS_LS, A, b, np.full(N, float(N))).set_reg(orc.REG_NORML1, lam=0.05)
    gam = 0.999 * N / (np.sum(A * A, axis=1) * N) * np.linspace(0.7, 1.3, N)
    x0 = np.full(d, 0.1)
    ref = orc.FinitoState(prob, x0, gam)                      # the whole problem, sequential loop over every batch
    batches = static_batches(N, 2 * B * world) * 2            # two epochs; every batch starts at a multiple of B·world rows
    mine = set(interleaved_rows(N, B, world, rank).tolist())  # 0-based global rows of this rank
    s = {i: ref.s[i].copy() for i in mine}                    # my table rows after init
    av, z, hat = ref.av.copy(), ref.z.copy(), ref.hat_gamma
    for batch in batches:
        part = np.zeros(d)
        for i1 in batch:
            i = int(i1) - 1
            if i in mine:
                t = z - (gam[i] / N) * prob.gradient(i, z)[0]
                part += (t - s[i]) * (hat / gam[i])
                s[i] = t
        tt = torch.from_numpy(part)
        dist.all_reduce(tt)                                   # the ranks' sums, then the closing update on every rank
        av = av + tt.numpy()
        z = prob.prox(av, hat)
    ref.steps(batches)
    err = max(np.abs(z - ref.z).max() / np.abs(ref.z).max(), np.abs(av - ref.av).max() / np.abs(ref.av).max(),
              max(np.abs(s[i] - ref.s[i]).max() for i in mine) / np.abs(ref.s).max())
    gathered = [None] * world
    dist.all_gather_object(gathered, float(err))
    if rank == 0:
        q.put(max(gathered))
    dist.destroy_process_group()


def test_interleaved_minibatch_epoch_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_minibatch_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    err = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert err < 1e-12
