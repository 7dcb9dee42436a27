"""debug: sample-sharded exchange over NCCL on a small input, with stage prints (torchrun)"""
import faulthandler, os, sys, time
faulthandler.enable()
faulthandler.dump_traceback_later(60, exit=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import simka_b200
from simka_b200 import dist as sdist
from helpers import synthetic_samples, oracle_run, assert_stats_equal
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
def P(*a): print(f"[r{rank}]", *a, flush=True)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
P("init ok")
samples = synthetic_samples(6, 3000)
N = len(samples)
cc = simka_b200.SimkaGpu(N, 31, 2, simple_dist=True, complex_dist=True, device=local, expected_kmers_per_sample=3000 * 120)
mc = simka_b200.SimkaGpu(N, 31, 2, simple_dist=True, complex_dist=True, device=local)
for s, (seqs, off) in enumerate(samples):
    if sdist.sample_owner(s, world) != rank: continue
    p, b, nw = simka_b200.pack_reads(seqs, off)
    cc.push_reads(s, p, b, nw, len(off) - 1); cc.finish_sample(s)
P("counted")
t = torch.ones(4, device=dev); dist.all_reduce(t); P("allreduce ok", t[0].item())
x = torch.arange(4, device=dev, dtype=torch.int64); y = torch.empty(4, device=dev, dtype=torch.int64)
dist.all_to_all_single(y, x, [2, 2], [2, 2]); torch.cuda.synchronize(); P("a2a ok", y.tolist())
st = sdist.exchange_and_merge(cc, mc, N, rank, world, dev)
P("merged")
st = sdist.allreduce_stats(st)
P("reduced")
if rank == 0:
    run = oracle_run(samples, 31, 2)
    assert_stats_equal(st, run)
    P("PARITY OK")
dist.barrier(); dist.destroy_process_group()
