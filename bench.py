#!/usr/bin/env python
"""bench.py -- headline benchmark of the Simka hot path on B200.

Metric (BASELINE.json): 31-mers/sec into the N x N distance matrices.  One "step" = the whole hot
path over one batch of synthetic input: for every sample canonical k-mer extraction -> bucket
partition -> exact counting -> abundance filter, then the multi-sample merge + N x N accumulator
update and the read-out of the accumulators.  Workload at 1 GPU = BASELINE.json configs[1]:
32 samples x 1M reads x 150 bp, k=31, abundance-min 2, simple + complex distances.

  value  : reads already 2-bit packed and resident in HBM when the timed region starts
  e2e    : the same through the C ABI with pinned HOST buffers (H2D of the packed reads and D2H of
           the accumulators inside the timed region, copy of sample s+1 overlapped with sample s)
  --gpus N>1 (torchrun): samples sharded over the ranks for counting, hash ranges sharded for the merge,
           solid k-mer lists exchanged with one NCCL all-to-all, two tiny all-reduces; same total
           workload => "scaling": "strong"
  --impl reference : the CPU restatement of the reference algorithm (oracle/, OpenMP, all host cores)
           on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=int, default=32)
    ap.add_argument("--reads", type=int, default=1_000_000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--kmer", type=int, default=31)
    ap.add_argument("--abundance-min", type=int, default=2)
    ap.add_argument("--ref-reads", type=int, default=400_000, help="reads/sample of the bounded CPU sample (about 10-20 s of CPU work per pass)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--genomes", type=int, default=0, help="number of 50 kb genomes of the generator (0 = 256 per 32 samples)")
    return ap.parse_args()


def workload_name(a, reads=None):
    return (f"synthetic {a.samples} samples x {reads or a.reads} reads x {a.read_len} bp, k={a.kmer}, "
            f"abundance-min={a.abundance_min}, simple+complex distances")


def n_genomes_for(samples, override=0):
    if override:
        return override
    return 256 if samples <= 32 else 256 * ((samples + 31) // 32)  # "scale G with N" 


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference algorithm, all host threads, bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_port_rate(a, reads, repeats=1):
    """-> (k-mers/s, cores, seconds per pass)"""
    from oracle import oracle
    from simka_b200 import synth
    genomes = synth.make_genomes(n_genomes_for(a.samples, a.genomes), 50_000)
    datas, offs, sample_off, base, nreads = [], [np.zeros(1, np.uint64)], [0], 0, 0
    for s in range(a.samples):
        seqs, off = synth.make_sample(genomes, s, reads, a.read_len)
        datas.append(seqs)
        offs.append(off[1:] + np.uint64(base))
        base += len(seqs)
        nreads += reads
        sample_off.append(nreads)
    seqs = np.concatenate(datas)
    ro = np.concatenate(offs)
    so = np.array(sample_off, np.uint64)
    kmers = a.samples * reads * (a.read_len - a.kmer + 1)
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        run = oracle.Run(seqs, ro, so, a.kmer, abundance_min=a.abundance_min, simple=True, complex_=True, nthreads=0)
        run.merge()
        for m in range(len(run.matrix_names())):
            run.matrix(m)
        dt = time.perf_counter() - t0
        run.close()
        best = dt if best is None else min(best, dt)
    return kmers / best, os.cpu_count(), best


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import oracle
    oracle.build()
    times = []
    kmers = a.samples * a.ref_reads * (a.read_len - a.kmer + 1)
    # build inputs once, time W + K passes
    from simka_b200 import synth
    genomes = synth.make_genomes(n_genomes_for(a.samples, a.genomes), 50_000)
    datas, offs, sample_off, base, nreads = [], [np.zeros(1, np.uint64)], [0], 0, 0
    for s in range(a.samples):
        seqs, off = synth.make_sample(genomes, s, a.ref_reads, a.read_len)
        datas.append(seqs)
        offs.append(off[1:] + np.uint64(base))
        base += len(seqs)
        nreads += a.ref_reads
        sample_off.append(nreads)
    seqs, ro, so = np.concatenate(datas), np.concatenate(offs), np.array(sample_off, np.uint64)
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        run = oracle.Run(seqs, ro, so, a.kmer, abundance_min=a.abundance_min, simple=True, complex_=True, nthreads=0)
        run.merge()
        for m in range(len(run.matrix_names())):
            run.matrix(m)
        run.close()
        if it >= a.warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    value = kmers * a.steps / total
    sample = f"{a.samples} samples x {a.ref_reads} reads x {a.read_len} bp per step (same generator, reads/sample scaled down)"
    line = {
        "impl": "reference", "metric": "31-mers/sec into NxN distance matrix", "value": value, "unit": "kmers/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload_name(a), "cpu_sample": sample},
        "cpu_baseline": {"value": value, "unit": "kmers/s", "cores": os.cpu_count(), "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md)
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        if not shutil.which("nvidia-smi"):
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, reasons = [], 0, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            v = d.get("hbm_gbs") or d.get("hbm_gbs_measured")
            if v:
                return float(v), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent)"


def run_ours(a):
    import torch
    import torch.distributed as dist
    import simka_b200
    from simka_b200 import build as sbuild
    from simka_b200 import dist as sdist
    from simka_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; simka_b200 has no CPU path")
    sbuild.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    # ---- synthetic input: rank r owns samples s with s % world == r (sample-sharded counting) ----------
    N, M, L, k = a.samples, a.reads, a.read_len, a.kmer
    genomes = synth.make_genomes(n_genomes_for(N, a.genomes), 50_000)
    t_gen = time.perf_counter()
    mine = [s for s in range(N) if sdist.sample_owner(s, world) == rank]
    host, devbuf, nwords = {}, {}, {}

    def gen(s):  # numpy and the packer release the GIL: the samples are generated on a few threads
        seqs, off = synth.make_sample(genomes, s, M, L)
        return s, simka_b200.pack_reads(seqs, off, nthreads=1, pinned=True)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max(1, min(8, (os.cpu_count() or 8) // max(1, world)))) as ex:
        for s, (hp, hb, nw) in ex.map(gen, mine):
            host[s] = (hp, hb)
            nwords[s] = nw
            devbuf[s] = (torch.from_numpy(np.asarray(hp).view(np.int64)).to(dev), torch.from_numpy(np.asarray(hb).view(np.int32)).to(dev))
    if world > 1:
        dist.barrier()
    gen_s = time.perf_counter() - t_gen
    kmers_per_step = N * M * (L - k + 1)
    h2d = sum((nw + 1) * 12 for nw in nwords.values())

    exp = M * (L - k + 1)
    ctx = simka_b200.SimkaGpu(N, k, a.abundance_min, simple_dist=True, complex_dist=True, device=local, expected_kmers_per_sample=exp)
    mctx = simka_b200.SimkaGpu(N, k, a.abundance_min, simple_dist=True, complex_dist=True, device=local) if world > 1 else None
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    d2h_bytes = [0]

    def finish_run():
        if world > 1:
            mctx.reset()
            st = sdist.exchange_and_merge(ctx, mctx, N, rank, world, dev)
        else:
            ctx.merge_and_accumulate()
            st = ctx.stats()
        d2h_bytes[0] = sum(getattr(st, n).nbytes for n in st.U64_SYM + st.U64_SQ) + st.kullback_leibler.nbytes
        if world > 1:
            st = sdist.allreduce_stats(st)
        return st

    def step_resident():
        ctx.reset()
        for s in mine:
            dp, db = devbuf[s]
            ctx.push_reads_device(s, dp.data_ptr(), db.data_ptr(), nwords[s], M)
            ctx.finish_sample(s)
        return finish_run()

    def step_e2e():
        ctx.reset()
        if mine:
            ctx.push_reads(mine[0], host[mine[0]][0], host[mine[0]][1], nwords[mine[0]], M)
        for i, s in enumerate(mine):
            if i + 1 < len(mine):
                s2 = mine[i + 1]
                ctx.push_reads(s2, host[s2][0], host[s2][1], nwords[s2], M)
            ctx.finish_sample(s)
        return finish_run()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tims = {}
        launches = 0
        e0.record(ext)
        st = None
        for _ in range(steps):
            st = fn()
            t = ctx.timings()
            if mctx is not None:  # stage 2 ran in the merge context
                tm = mctx.timings()
                for kk in ("merge_ms", "distinct_kmers", "pairs", "kernel_launches"):
                    t[kk] += tm[kk]
                t["total_ms"] += tm["merge_ms"]
            launches += t["kernel_launches"]
            for kk, v in t.items():
                tims[kk] = tims.get(kk, 0) + v
        e1.record(ext)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, tims, launches, st

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_res, tims, launches, st = timed(step_resident, a.steps, a.warmup)
    ms_e2e, _, _, st2 = timed(step_e2e, a.steps, max(1, a.warmup - 2))
    clocks = sampler.stop() if rank == 0 else {}
    # sanity: both paths must produce the same accumulators
    assert np.array_equal(st.nb_shared, st2.nb_shared) and np.array_equal(st.braycurtis_num, st2.braycurtis_num)
    if world > 1:  # whole-job byte and launch counts
        t = torch.tensor([h2d, launches, d2h_bytes[0]], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d, launches, d2h_bytes[0] = (int(x) for x in t.tolist())

    if rank == 0:
        steps = a.steps
        per = {kk: v / steps for kk, v in tims.items()}
        Ks, D, U = per["kmers_in"], per["entries"], per["distinct_kmers"]
        words = sum(nwords.values())
        # algorithmic bytes per launch set (summed over the samples of one step); DESIGN.md section 4
        superk = per.get("superk_samples", 0) > 0  # SIMKA_B200_SUPERK=1: experimental super-k-mer front end of stage 1
        if superk:
            # the scan reads the packed reads and the expansion writes the keys: together the algorithmic bytes of the k-mer
            # extraction; the record list in between (about 2 B per k-mer each way) is implementation traffic
            alg = {
                "k_superk_scan": words * 12,
                "k_superk_bins(+scan)": 0,
                "k_superk_expand": 8 * Ks,
                "k_bin_count": 8 * Ks + 12 * D,
                "k_merge": 12 * D + 8 * U,
            }
            tms = {
                "k_superk_scan": per["extract_scatter_ms"], "k_superk_bins(+scan)": per["part_count_ms"],
                "k_superk_expand": per["part_scatter_ms"], "k_bin_count": per["bin_count_ms"], "k_merge": per["merge_ms"],
            }
        else:
            alg = {
                "k_extract<count>": words * 12,
                "k_extract<scatter>": words * 12 + 8 * Ks,
                "k_partition2<count>": 8 * Ks,
                "k_partition2<scatter>": 16 * Ks,
                "k_bin_count": 8 * Ks + 12 * D,
                "k_merge": 12 * D + 8 * U,
            }
            tms = {
                "k_extract<count>": per["extract_count_ms"], "k_extract<scatter>": per["extract_scatter_ms"],
                "k_partition2<count>": per["part_count_ms"], "k_partition2<scatter>": per["part_scatter_ms"],
                "k_bin_count": per["bin_count_ms"], "k_merge": per["merge_ms"],
            }
        peak, peak_src = hbm_peak()
        kernels = {n: {"ms_per_step": tms[n], "alg_bytes_per_step": alg[n],
                       "achieved_gbs": (alg[n] / (tms[n] * 1e-3) / 1e9) if tms[n] > 0 else None} for n in alg}
        dom = max(tms, key=lambda n: tms[n])
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(dom)
            except Exception:
                traffic = None
        achieved = kernels[dom]["achieved_gbs"] or 0.0
        line = {
            "metric": "31-mers/sec into NxN distance matrix", "value": kmers_per_step * steps / (ms_res * 1e-3),
            "unit": "kmers/s", "n_gpus": world, "steps": steps, "warmup": a.warmup, "ms_per_step": ms_res / steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload_name(a), "sharding": (f"samples over {world} ranks for counting (rank r owns samples s % {world} == r), hash ranges over the ranks for the merge; "
                                    "solid (k-mer,count) lists exchanged with NCCL all-to-all, 2 tiny all-reduces; kernels/roofline are rank 0's share") if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (packed reads 57 MB/sample, 0.96 GB of keys per sample streamed)",
                       "generator": "BASELINE.md section 3 (G=256 x 50 kb, Dirichlet(0.3), 1% substitutions, 50% revcomp)",
                       "input_build_s": round(gen_s, 1),
                       "stage1": ("super-k-mer front end (SIMKA_B200_SUPERK=1, experimental), %d fallbacks" % per.get("superk_fallbacks", 0)) if superk
                                 else "k-mer partition (default)"},
            "e2e": {"value": kmers_per_step * steps / (ms_e2e * 1e-3), "unit": "kmers/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_bytes[0]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "share_of_step": tms[dom] / (ms_res / steps)},
            "kernels": kernels,
            "counts": {"kmers_in_per_step": Ks, "solid_entries": D, "distinct_kmers": U, "pairs": per["pairs"]},
        }
        if world == 1 and not a.no_cpu_baseline:
            from oracle import oracle
            oracle.build()
            rate, cores, secs = cpu_port_rate(a, a.ref_reads)
            line["cpu_baseline"] = {"value": rate, "unit": "kmers/s", "cores": cores, "kind": "port",
                                    "sample": f"{N} samples x {a.ref_reads} reads x {L} bp (same generator), one pass of {secs:.1f} s"}
        print(json.dumps(line), flush=True)
    ctx.close()
    if mctx is not None:
        mctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    if os.environ.get("SIMKA_BENCH_WATCHDOG"):  # dump all stacks and exit if the run takes longer than this many seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["SIMKA_BENCH_WATCHDOG"]), exit=True)
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
