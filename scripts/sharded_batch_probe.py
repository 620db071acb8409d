"""Minibatch sweeps on row shards (one process per GPU, torchrun): the persistent kernel with the in-kernel rank exchange against
one pass + tail kernel per batch (CIAO_BATCH_PER_LAUNCH=1).  C2 shape: logistic N = 2^20 x 1024, rows split over the ranks.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port 29577 scripts/sharded_batch_probe.py
"""
import os, sys, json, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); import ciao_pkg; ciao_pkg.load()
from ciaoalgorithms_jl_b200 import _lib as L
from ciaoalgorithms_jl_b200.engine import Engine
from ciaoalgorithms_jl_b200.sampling import BatchSweeper, HostRNG, csr, interleaved_rows, shard_rows

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, d = 1 << 20, 1024
lo, hi = shard_rows(N, world, rank)
res = {}
BLOCK = 256
for per_launch, layout in (("0", "interleaved"), ("0", "contiguous"), ("1", "contiguous")):
    os.environ["CIAO_BATCH_PER_LAUNCH"] = per_launch        # read by ciao_create
    e = Engine(local)
    if layout == "interleaved":                             # blocks of 256 rows dealt round-robin: every batch is spread over all ranks
        e.set_row_interleave(BLOCK, rank, world)
        e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0, row0=rank * BLOCK, n_rows=len(interleaved_rows(N, BLOCK, world, rank)))
    else:
        e.gen_synthetic(L.SYNTH_LOGISTIC, N, d, 0x5EED0002, scale=1.0, row0=lo, n_rows=hi - lo)
    e.set_reg(L.REG_NORML1, 1.0 / N)
    hs = [None] * world
    dist.all_gather_object(hs, e.comm_p2p_handle())
    e.comm_p2p_attach(rank, world, hs)
    Lmax = 0.25 * e.max_row_sqnorm(); gam = np.full(N, 0.999 * N / Lmax); hat = 1 / np.sum(1 / gam)
    for r in (4096, 4096 * world, 65536):
        sw = BatchSweeper(N, r, 2, HostRNG(1)); idx, bp = csr(sw.take(sw.d))
        e.finito_init(np.ones(d), gam, hat)
        e.finito_steps(idx, bp); dist.barrier(); e.finito_steps(idx, bp); tf = e.last_timing().last_seq_ms
        e.lfinito_init(np.ones(d), gam, hat)
        o = np.arange(1, sw.d + 1)
        e.lfinito_outer(o, r); dist.barrier(); e.lfinito_outer(o, r); tl = e.last_timing().last_seq_ms
        z = e.get_vec(L.VEC_Z)
        blobs = [None] * world; dist.all_gather_object(blobs, z.tobytes())
        t = torch.tensor([tf, tl], device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); tf, tl = t.tolist()
        key = f"{layout}_{'per_batch' if per_launch == '1' else 'fused'}_batch{r}"
        res[key] = {"finito_us_per_batch": 1e3 * tf / sw.d, "lfinito_us_per_batch": 1e3 * tl / sw.d, "finito_epochs_per_s": 1e3 / tf,
                    "lfinito_sweeps_per_s": 1e3 / tl, "z_bitwise_equal_across_ranks": all(b == blobs[0] for b in blobs)}
        if rank == 0:
            print(f"[{world} GPUs] {key}: finito {1e3 * tf / sw.d:.1f} us/batch ({1e3 / tf:.1f} epochs/s), lfinito {1e3 * tl / sw.d:.1f} us/batch ({1e3 / tl:.1f} sweeps/s), "
                  f"same z on all ranks: {res[key]['z_bitwise_equal_across_ranks']}", flush=True)
    e.close()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"world": world, "N": N, "d": d, "results": res}, open(f"gpurun_out/sharded_batch_n{world}.json", "w"), indent=1)
dist.destroy_process_group()
