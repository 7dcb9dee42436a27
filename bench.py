#!/usr/bin/env python
"""bench.py -- headline benchmark of the Simka hot path on B200.

Metric (BASELINE.json): 31-mers/sec into the N x N distance matrices.  One "step" = the whole hot
path over one batch of synthetic input: for every sample canonical k-mer extraction -> bucket
partition -> exact counting -> abundance filter, then the multi-sample merge + N x N accumulator
update and the read-out of the accumulators.

  --config 2 (default)  BASELINE.json configs[1]: 32 samples x 1M reads x 150 bp, k=31, abundance-min 2, 1 GPU
  --config 3            configs[2]: 128 x 5M, k=31 (meant for --gpus 8 under torchrun)
  --config 4            configs[3]: 512 x 1M, k=21, 1 GPU (the 512 x 512 accumulator update)
  --config 5            configs[4]: SimkaMin, 128 sketches x 1M k-mers, sketch x sketch Jaccard / Bray-Curtis on 1 GPU
  (--samples / --reads / --kmer override the configuration's sizes)

  value  : reads already 2-bit packed and resident in HBM when the timed region starts
  e2e    : the same through the C ABI with pinned HOST buffers (H2D of the packed reads and D2H of
           the accumulators inside the timed region, copy of sample s+1 overlapped with sample s)
  --gpus N>1 (torchrun): samples sharded over the ranks for counting, hash ranges sharded for the merge,
           solid k-mer lists exchanged over NCCL, two tiny all-reduces; same total workload => "scaling": "strong"
  --impl reference : the REFERENCE BINARIES (oracle/_ref/bin/simka, built from /root/reference by oracle/build_ref.sh)
           on FASTA files in /dev/shm, all host cores, on a bounded sample of the same workload; the oracle port
           (oracle/, OpenMP) when the binaries are absent.
  --verify : after the timing, every accumulator of the step is compared with the CPU oracle on the same reads
           (integers bit-exact, Jensen-Shannon sums within 1e-10 relative); N = 1 only.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent
CONFIGS = {2: (32, 1_000_000, 31), 3: (128, 5_000_000, 31), 4: (512, 1_000_000, 21), 5: (128, 1_000_000, 21)}
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "bin")


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--samples", type=int, default=0)
    ap.add_argument("--reads", type=int, default=0, help="reads per sample (config 5: k-mers per sketch)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--kmer", type=int, default=0)
    ap.add_argument("--abundance-min", type=int, default=2)
    ap.add_argument("--ref-reads", type=int, default=0, help="reads/sample of the bounded CPU sample (0 = sized from a calibration run)")
    ap.add_argument("--ref-budget-s", type=float, default=300.0, help="wall-clock target of the whole --impl reference run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--verify", action="store_true", help="compare all accumulators with the CPU oracle after the timing (N = 1)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end pass (0 = --steps)")
    ap.add_argument("--genomes", type=int, default=0, help="number of 50 kb genomes of the generator (0 = 256 per 32 samples)")
    ap.add_argument("--out", default="", help="also write the JSON line to this file")
    a = ap.parse_args()
    n, m, k = CONFIGS[a.config]
    a.samples = a.samples or n
    a.reads = a.reads or m
    a.kmer = a.kmer or k
    return a


def workload_name(a):
    if a.config == 5:
        return f"SimkaMin distance: {a.samples} sketches x {a.reads} k-mers, all {a.samples * (a.samples - 1) // 2} pairs, Jaccard + Bray-Curtis"
    return (f"synthetic {a.samples} samples x {a.reads} reads x {a.read_len} bp, k={a.kmer}, "
            f"abundance-min={a.abundance_min}, simple+complex distances")


def config_dict(a):
    """identical for both arms (the driver compares them)"""
    if a.config == 5:
        return {"workload": workload_name(a), "generator": "BASELINE.md section 3 sketches (pool of 4 s hashes, s drawn per sketch, counts 1..19)"}
    return {"workload": workload_name(a),
            "generator": f"BASELINE.md section 3 (G={n_genomes_for(a.samples, a.genomes)} x 50 kb, Dirichlet(0.3), 1% substitutions, 50% revcomp)",
            "l2": "inputs larger than L2 (packed reads 57 MB per 1M-read sample, 0.96 GB of keys per sample streamed)"}


def n_genomes_for(samples, override=0):
    if override:
        return override
    return 256 if samples <= 32 else 256 * ((samples + 31) // 32)  # "scale G with N" 


def emit(a, line):
    s = json.dumps(line)
    print(s, flush=True)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        with open(a.out, "w") as f:
            f.write(s + "\n")


# ----------------------------------------------------------------------------------------------
# CPU arms
# ----------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def concat_samples(a, reads):
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
    return np.concatenate(datas), np.concatenate(offs), np.array(sample_off, np.uint64)


def oracle_pass(a, seqs, ro, so):
    """one pass of the oracle port (all host threads; torchrun's OMP_NUM_THREADS=1 is ignored on purpose) -> seconds"""
    from oracle import oracle
    t0 = time.perf_counter()
    run = oracle.Run(seqs, ro, so, a.kmer, abundance_min=a.abundance_min, simple=True, complex_=True, nthreads=host_cores())
    run.merge()
    for m in range(len(run.matrix_names())):
        run.matrix(m)
    dt = time.perf_counter() - t0
    run.close()
    return dt


class RefBinaryRun:
    """The reference's own CPU implementation: `simka` (which starts one simkaCount per sample and one simkaMerge per
    partition) on FASTA files of the same generator, temp dir on /dev/shm (the reference is disk-bound by design)."""

    def __init__(self, a):
        self.a = a
        base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
        self.dir = tempfile.mkdtemp(prefix="simka_bench_ref_", dir=base)
        self.where = "/dev/shm" if base else "tmp"
        self.cores = host_cores()
        self.reads = 0

    @staticmethod
    def available():
        return all(os.access(os.path.join(REF_BIN, b), os.X_OK) for b in ("simka", "simkaCount", "simkaMerge", "simkaCountProcess"))

    def write_inputs(self, reads):
        from simka_b200 import synth
        a = self.a
        ind = os.path.join(self.dir, f"in_{reads}")
        if os.path.isdir(ind):
            shutil.rmtree(ind)
        os.makedirs(ind)
        genomes = synth.make_genomes(n_genomes_for(a.samples, a.genomes), 50_000)
        L = a.read_len
        with open(os.path.join(ind, "input.txt"), "w") as f:
            for s in range(a.samples):
                seqs, _ = synth.make_sample(genomes, s, reads, L)
                rec = np.empty((reads, L + 4), np.uint8)  # ">r\n" + bases + "\n"
                rec[:, 0], rec[:, 1], rec[:, 2], rec[:, L + 3] = ord(">"), ord("r"), 10, 10
                rec[:, 3:L + 3] = seqs.reshape(reads, L)
                rec.tofile(os.path.join(ind, f"S{s}.fasta"))
                f.write(f"S{s}: S{s}.fasta\n")
        self.reads = reads
        return ind

    def run_once(self, ind):
        a = self.a
        out, tmp = os.path.join(self.dir, "out"), os.path.join(self.dir, "tmp")
        for d in (out, tmp):
            shutil.rmtree(d, ignore_errors=True)
        os.makedirs(tmp)
        try:
            ram_mb = os.sysconf("SC_PHYS_PAGES") * os.sysconf("SC_PAGE_SIZE") // (1 << 20)
        except Exception:
            ram_mb = 64000
        mem = int(max(2000, min(1000 * self.cores, ram_mb // 2)))
        cmd = [os.path.join(REF_BIN, "simka"), "-in", os.path.join(ind, "input.txt"), "-out", out, "-out-tmp", tmp, "-kmer-size", str(a.kmer),
               "-abundance-min", str(a.abundance_min), "-simple-dist", "-complex-dist", "-nb-cores", str(self.cores), "-max-memory", str(mem)]
        t0 = time.perf_counter()
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        dt = time.perf_counter() - t0
        if r.returncode != 0 or not os.path.exists(os.path.join(out, "mat_abundance_braycurtis.csv.gz")):
            raise RuntimeError("reference simka failed: " + r.stdout[-600:])
        return dt

    def kmers(self):
        a = self.a
        return a.samples * self.reads * (a.read_len - a.kmer + 1)

    def sample_text(self, secs):
        a = self.a
        return (f"reference binaries (simka -nb-cores {self.cores}, FASTA + temp dir on {self.where}): {a.samples} samples x {self.reads} reads x "
                f"{a.read_len} bp of the same generator, {secs:.1f} s per run")

    def close(self):
        shutil.rmtree(self.dir, ignore_errors=True)


def ref_sample_size(ref, a, per_run_s):
    """reads per sample such that one run of the reference takes about per_run_s.  The reference pays seconds of process start-up and
    polling per run whatever the input size, so the size is fitted on TWO calibration runs (fixed cost + cost per read) -- a single
    small run would size the sample far too small and understate the reference's rate.  -> (reads, seconds spent calibrating)"""
    c1, c2 = min(a.reads, 8_000), min(a.reads, 32_000)
    t1 = ref.run_once(ref.write_inputs(c1))
    if c2 == c1:
        return c1, t1
    t2 = ref.run_once(ref.write_inputs(c2))
    per_read = max((t2 - t1) / (c2 - c1), 1e-7)
    fixed = max(0.0, t1 - per_read * c1)
    reads = int(min(a.reads, max(c2, (per_run_s - fixed) / per_read)))
    return max(1000, reads // 1000 * 1000), t1 + t2


def cpu_baseline_once(a, target_s=15.0):
    """bounded CPU baseline for our arm's line (N = 1): one run of the reference binaries (else the oracle port) of about target_s"""
    if RefBinaryRun.available():
        ref = RefBinaryRun(a)
        try:
            reads, _ = ref_sample_size(ref, a, target_s)
            secs = ref.run_once(ref.write_inputs(reads))
            return {"value": ref.kmers() / secs, "unit": "kmers/s", "cores": ref.cores, "kind": "reference", "sample": ref.sample_text(secs)}
        finally:
            ref.close()
    from oracle import oracle
    oracle.build()
    reads = min(a.reads, 400_000 * 32 // max(32, a.samples))
    secs = oracle_pass(a, *concat_samples(a, reads))
    return {"value": a.samples * reads * (a.read_len - a.kmer + 1) / secs, "unit": "kmers/s", "cores": host_cores(), "kind": "port",
            "sample": f"oracle port (OpenMP): {a.samples} samples x {reads} reads x {a.read_len} bp (same generator), one pass of {secs:.1f} s"}


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    if a.config == 5:
        return run_reference_simkamin(a)
    nrun = a.warmup + a.steps
    times = []
    if RefBinaryRun.available():
        ref = RefBinaryRun(a)
        try:
            reads = a.ref_reads
            if not reads:  # size the sample so that warmup + steps runs fit the budget
                reads, t_cal = ref_sample_size(ref, a, max(2.0, (a.ref_budget_s - 15.0) / nrun * 0.9))
            ind = ref.write_inputs(reads)
            for it in range(nrun):
                dt = ref.run_once(ind)
                if it >= a.warmup:
                    times.append(dt)
            kmers, cores, kind = ref.kmers(), ref.cores, "reference"
            sample = ref.sample_text(sum(times) / len(times))
        finally:
            ref.close()
    else:
        from oracle import oracle
        oracle.build()
        reads = a.ref_reads or min(a.reads, 400_000 * 32 // max(32, a.samples))
        seqs, ro, so = concat_samples(a, reads)
        for it in range(nrun):
            dt = oracle_pass(a, seqs, ro, so)
            if it >= a.warmup:
                times.append(dt)
        kmers, cores, kind = a.samples * reads * (a.read_len - a.kmer + 1), host_cores(), "port"
        sample = f"oracle port (OpenMP): {a.samples} samples x {reads} reads x {a.read_len} bp per step (same generator, reads/sample scaled down)"
    total = sum(times)
    value = kmers * a.steps / total
    emit(a, {
        "impl": "reference", "metric": "31-mers/sec into NxN distance matrix", "value": value, "unit": "kmers/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(a),
        "cpu_baseline": {"value": value, "unit": "kmers/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "kmers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    })


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


def measured_traffic(kernel, kmers_this_rank):
    """DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of `kernel` per launch set, from the ncu --set full capture of
    the shipped build: profiles/traffic.json holds bytes PER INPUT K-MER of the capture, scaled here by the k-mers this rank
    processed (so it follows the workload, not a constant); null when the file does not list the kernel."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        d = json.load(open(tp))
        per = d.get("bytes_per_kmer", {}).get(kernel)
        return None if per is None else float(per) * kmers_this_rank
    except Exception:
        return None


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
    if rank == 0:
        sbuild.build()
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()
    dev = torch.device("cuda", local)
    if a.config == 5:
        return run_simkamin(a, torch)

    # ---- synthetic input: rank r owns samples s with s % world == r (sample-sharded counting) ----------
    N, M, L, k = a.samples, a.reads, a.read_len, a.kmer
    genomes = synth.make_genomes(n_genomes_for(N, a.genomes), 50_000)
    t_gen = time.perf_counter()
    mine = [s for s in range(N) if sdist.sample_owner(s, world) == rank]
    host, devbuf, nwords, ascii_keep = {}, {}, {}, {}
    keep_ascii = a.verify and world == 1

    def gen(s):  # numpy and the packer release the GIL: the samples are generated on a few threads
        seqs, off = synth.make_sample(genomes, s, M, L)
        return s, simka_b200.pack_reads(seqs, off, nthreads=1, pinned=True), (seqs if keep_ascii else None)

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max(1, min(8, host_cores() // max(1, world)))) as ex:
        for s, (hp, hb, nw), seqs in ex.map(gen, mine):
            host[s] = (hp, hb)
            nwords[s] = nw
            if keep_ascii:
                ascii_keep[s] = seqs
            devbuf[s] = (torch.from_numpy(np.asarray(hp).view(np.int64)).to(dev), torch.from_numpy(np.asarray(hb).view(np.int32)).to(dev))
    if world > 1:
        dist.barrier()
    gen_s = time.perf_counter() - t_gen
    kmers_per_step = N * M * (L - k + 1)
    h2d = sum((nw + 1) * 12 for nw in nwords.values())

    exp = M * (L - k + 1)
    ctx = simka_b200.SimkaGpu(N, k, a.abundance_min, simple_dist=True, complex_dist=True, device=local, expected_kmers_per_sample=exp)
    mctx = simka_b200.SimkaGpu(N, k, a.abundance_min, simple_dist=True, complex_dist=True, device=local) if (world > 1 and os.environ.get("SIMKA_BENCH_EXCHANGE", "native") != "native") else None
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    d2h_bytes = [0]

    native = world > 1 and os.environ.get("SIMKA_BENCH_EXCHANGE", "native") == "native"  # "torch": the round-1 exchange in simka_b200/dist.py
    if native:
        sdist.native_comm_init(ctx, rank, world, dev)

    def finish_run():
        if native:  # totals all-reduce, list exchange, merge, accumulator all-reduce: all inside the library (NCCL)
            st = sdist.exchange_and_merge_native(ctx)
        elif world > 1:
            mctx.reset()
            st = sdist.exchange_and_merge(ctx, mctx, N, rank, world, dev)
        else:
            ctx.merge_and_accumulate()
            st = ctx.stats()
        d2h_bytes[0] = sum(getattr(st, n).nbytes for n in st.U64_SYM + st.U64_SQ) + st.kullback_leibler.nbytes
        if world > 1 and not native:
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
            if native:  # (ctx.timings() already reports stage 2 of the internal merge context; pairs / distinct k-mers are global there)
                xt = ctx.exchange_timings()
                for kk in ("totals_sizes_ms", "exchange_ms", "allreduce_ms"):
                    t["x_" + kk] = xt[kk]
            elif mctx is not None:  # stage 2 ran in the merge context
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
    e2e_steps = a.e2e_steps or a.steps
    ms_e2e, _, _, st2 = timed(step_e2e, e2e_steps, max(1, min(a.warmup, 3) - 2))
    clocks = sampler.stop() if rank == 0 else {}
    # sanity: both paths must produce the same accumulators
    assert np.array_equal(st.nb_shared, st2.nb_shared) and np.array_equal(st.braycurtis_num, st2.braycurtis_num)
    if world > 1:  # whole-job byte and launch counts
        t = torch.tensor([h2d, launches, d2h_bytes[0]], device=dev, dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        h2d, launches, d2h_bytes[0] = (int(x) for x in t.tolist())

    verify = None
    if keep_ascii and rank == 0:  # outside every timed region: ALL accumulators of the full-size step against the CPU oracle
        from oracle import oracle
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import assert_stats_equal
        oracle.build()
        t0 = time.perf_counter()
        seqs = np.concatenate([ascii_keep[s] for s in range(N)])
        ro = np.arange(N * M + 1, dtype=np.uint64) * np.uint64(L)
        so = np.arange(N + 1, dtype=np.uint64) * np.uint64(M)
        run = oracle.Run(seqs, ro, so, k, abundance_min=a.abundance_min, simple=True, complex_=True, nthreads=host_cores())
        run.merge()
        assert_stats_equal(st2, run, kl_rtol=1e-10)
        run.close()
        verify = {"against": "oracle port on the same reads", "checked": "all accumulators of the step (integers bit-exact, Jensen-Shannon sums rtol 1e-10) and the 21 matrices",
                  "oracle_s": round(time.perf_counter() - t0, 1), "result": "equal"}

    if rank == 0:
        steps = a.steps
        per = {kk: v / steps for kk, v in tims.items()}
        Ks, D, U = per["kmers_in"], per["entries"], per["distinct_kmers"]
        words = sum(nwords.values())
        # algorithmic bytes per launch set (summed over the samples of one step); DESIGN.md section 4
        superk = per.get("superk_samples", 0) > 0  # super-k-mer front end of stage 1
        if superk and os.environ.get("SIMKA_B200_SUPERK") == "2":
            # record path: SURVEY 8(d)'s per-unit figures (K1 = packed reads in + 8 B per k-mer out, K2 = 8 B per k-mer in + 12 B per
            # solid entry out) -- the keys themselves never cross HBM here, records of ~2.2 B per k-mer do (see "traffic")
            alg = {"k_superk_scan": words * 12 + 8 * Ks, "prefix_scans+k_superk_bins": 0, "k_sk2_scatter": 0, "k_sk2_count": 8 * Ks + 12 * D,
                   "k_merge": 12 * D + 8 * U}
            tms = {"k_superk_scan": per["extract_scatter_ms"], "prefix_scans+k_superk_bins": per["part_count_ms"],
                   "k_sk2_scatter": per["part_scatter_ms"], "k_sk2_count": per["bin_count_ms"], "k_merge": per["merge_ms"]}
        elif superk:
            alg = {"k_superk_scan": words * 12, "k_superk_bins(+scan)": 0, "k_superk_expand": 8 * Ks, "k_bin_count": 8 * Ks + 12 * D,
                   "k_merge": 12 * D + 8 * U}
            tms = {"k_superk_scan": per["extract_scatter_ms"], "k_superk_bins(+scan)": per["part_count_ms"],
                   "k_superk_expand": per["part_scatter_ms"], "k_bin_count": per["bin_count_ms"], "k_merge": per["merge_ms"]}
        else:
            alg = {"k_extract<count>": words * 12, "k_extract<scatter>": words * 12 + 8 * Ks, "k_partition2<count>": 8 * Ks,
                   "k_partition2<scatter>": 16 * Ks, "k_bin_count": 8 * Ks + 12 * D, "k_merge": 12 * D + 8 * U}
            tms = {"k_extract<count>": per["extract_count_ms"], "k_extract<scatter>": per["extract_scatter_ms"],
                   "k_partition2<count>": per["part_count_ms"], "k_partition2<scatter>": per["part_scatter_ms"],
                   "k_bin_count": per["bin_count_ms"], "k_merge": per["merge_ms"]}
        # a pass that did not run on this path (e.g. the exact K1 count pass) is not evidence: dropped
        for n in [n for n in tms if tms[n] <= 1e-2 * max(tms.values()) and n != "k_merge"]:
            tms.pop(n)
            alg.pop(n)
        peak, peak_src = hbm_peak()
        kernels = {n: {"ms_per_step": tms[n], "alg_bytes_per_step": alg[n], "achieved_gbs": (alg[n] / (tms[n] * 1e-3) / 1e9) if tms[n] > 0 else None} for n in alg}
        kernels["other (compaction, regrouping, scans, memsets)"] = {"ms_per_step": max(0.0, per["samples_ms"] + per["merge_ms"] - sum(tms.values()))}
        kernels["k_merge"]["pair_updates_per_s"] = per["pairs"] / (tms["k_merge"] * 1e-3) if (tms["k_merge"] > 0 and world == 1) else None
        dom = max(tms, key=lambda n: tms[n])
        achieved = kernels[dom]["achieved_gbs"] or 0.0
        step_ms = ms_res / steps
        # whole pipeline against the HBM roofline: SURVEY 8(d) end-to-end algorithmic bytes of THIS rank's share / step time
        pipe_alg = words * 12 + 16 * Ks + 12 * (Ks * 0.0 + D) + 24 * D + 16 * U  # (D0 is not measured: its 12 B/entry term uses D, a lower bound)
        line = {
            "metric": "31-mers/sec into NxN distance matrix", "value": kmers_per_step * steps / (ms_res * 1e-3),
            "unit": "kmers/s", "n_gpus": world, "steps": steps, "warmup": a.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config_dict(a),
            "notes": {"sharding": (f"samples over {world} ranks for counting (rank r owns samples s % {world} == r), hash ranges over the ranks for the merge; "
                                   "solid (k-mer,count) lists exchanged over NCCL, 2 tiny all-reduces; kernels/roofline are rank 0's share") if world > 1 else "single GPU",
                      "input_build_s": round(gen_s, 1),
                      "stage1": ("super-k-mer front end, %d fallbacks" % per.get("superk_fallbacks", 0)) if superk else "k-mer partition"},
            "e2e": {"value": kmers_per_step * e2e_steps / (ms_e2e * 1e-3), "unit": "kmers/s", "ms_per_step": ms_e2e / e2e_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_bytes[0]},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": measured_traffic(dom, Ks), "peak_source": peak_src,
                         "share_of_step": tms[dom] / step_ms,
                         "pipeline": {"alg_bytes_per_step": pipe_alg, "achieved": pipe_alg / (step_ms * 1e-3) / 1e9, "frac": pipe_alg / (step_ms * 1e-3) / 1e9 / peak}},
            "kernels": kernels,
            "counts": {"kmers_in_per_step": Ks, "solid_entries": D, "distinct_kmers": U, "pairs": per["pairs"]},
        }
        if world > 1 and "x_exchange_ms" in per:
            line["notes"]["exchange_rank0_ms_per_step"] = {"totals+sizes all-reduce": per["x_totals_sizes_ms"], "list exchange (ncclSend/ncclRecv)": per["x_exchange_ms"],
                                                           "accumulator all-reduce": per["x_allreduce_ms"], "implementation": "simka_gpu_exchange_and_merge (C++, NCCL)"}
        if verify:
            line["verify"] = verify
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_once(a)
        emit(a, line)
    ctx.close()
    if mctx is not None:
        mctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# config 5: SimkaMin sketch x sketch distances
# ----------------------------------------------------------------------------------------------
def run_simkamin(a, torch):
    import simka_b200
    from simka_b200 import synth
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n, s = a.samples, a.reads
    hashes, counts = synth.make_sketches(n, s)
    # the sketches live in pinned host memory (as the packed reads of the k-mer configurations do): the library's cudaMemcpy of a
    # pageable 1.5 GB would be staged through the driver's bounce buffer at a fifth of the PCIe rate
    if os.environ.get("SIMKA_BENCH_PIN", "1") != "0":
        pin_h, pin_c = torch.from_numpy(hashes).pin_memory(), torch.from_numpy(counts).pin_memory()
        hashes, counts = pin_h.numpy(), pin_c.numpy()
    pairs = n * (n - 1) // 2
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    kms, wall = [], []
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        jac, bc, ms = simka_b200.simkamin_distance(hashes, counts)  # host sketches in, host float matrices out
        w = time.perf_counter() - t0
        if os.environ.get("SIMKA_BENCH_VERBOSE"):
            print(f"step {it}: wall {w * 1e3:.1f} ms, kernel {ms:.1f} ms", file=sys.stderr, flush=True)
        if it >= a.warmup:
            kms.append(ms)
            wall.append(w)
    clocks = sampler.stop()
    # algorithmic bytes: 12 B (u64 hash + u32 count) per element the two-pointer merge of a pair consumes before it has seen s
    # distinct hashes (SURVEY 8d "per-pair streamed bytes"); counted exactly on the host from the sketches
    consumed, sampled = 0, 0
    for j in range(1, min(n, 17)):  # the sketches are i.i.d. draws from one pool: the mean over 16 pairs stands for all of them
        u = np.union1d(hashes[0], hashes[j])
        cut = u[min(s, len(u)) - 1]
        consumed += int(np.searchsorted(hashes[0], cut, "right")) + int(np.searchsorted(hashes[j], cut, "right"))
        sampled += 1
    alg = consumed * pairs // max(1, sampled) * 12
    k_ms, e_s = sum(kms) / len(kms), sum(wall) / len(wall)
    peak, peak_src = hbm_peak()
    line = {"metric": "SimkaMin sketch pairs/sec (Jaccard + Bray-Curtis)", "value": pairs / (k_ms * 1e-3), "unit": "pairs/s", "n_gpus": 1, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": k_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": config_dict(a),
            "e2e": {"value": pairs / e_s, "unit": "pairs/s", "ms_per_step": 1e3 * e_s, "h2d_bytes_per_step": int(hashes.nbytes + counts.nbytes), "d2h_bytes_per_step": 2 * n * n * 4},
            "gpu_launches": 2 * a.steps, "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_simkamin_tiled", "achieved": alg / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": alg / (k_ms * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src, "alg_bytes_per_step": alg}}
    if not a.no_cpu_baseline:
        from oracle import oracle
        oracle.build()
        m = min(32, n)
        t0 = time.perf_counter()
        oj, ob = oracle.simkamin_distance(hashes[:m], counts[:m])
        dt = time.perf_counter() - t0
        np.testing.assert_array_equal(jac[:m, :m], oj)
        np.testing.assert_array_equal(bc[:m, :m], ob)
        line["cpu_baseline"] = {"value": (m * (m - 1) // 2) / dt, "unit": "pairs/s", "cores": host_cores(), "kind": "port",
                                "sample": f"oracle port of SimkaMinDistance.hpp:190-279 on {m} of the sketches, {dt:.1f} s; GPU result bit-identical on that block"}
    emit(a, line)


def run_reference_simkamin(a):
    from oracle import oracle
    from simka_b200 import synth
    oracle.build()
    m = min(32, a.samples)
    hashes, counts = synth.make_sketches(m, a.reads)
    times = []
    for it in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        oracle.simkamin_distance(hashes, counts)
        if it >= a.warmup:
            times.append(time.perf_counter() - t0)
    value = (m * (m - 1) // 2) * a.steps / sum(times)
    sample = f"oracle port of SimkaMinDistance.hpp:190-279: {m} sketches x {a.reads} k-mers per step"
    emit(a, {"impl": "reference", "metric": "SimkaMin sketch pairs/sec (Jaccard + Bray-Curtis)", "value": value, "unit": "pairs/s", "n_gpus": a.gpus,
             "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / a.steps, "higher_is_better": True, "scaling": "strong",
             "vs_baseline": None, "dtype": "u64", "data": "synthetic", "config": config_dict(a),
             "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": host_cores(), "kind": "port", "sample": sample},
             "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})


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
