#!/usr/bin/env python
"""bench.py -- DRACH sites/sec of the MIL-inference hot path at num_iterations=1000.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--config 2|3|4|5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (default = BASELINE.json configs[2], the configuration the metric is quoted on): synthetic 1M DRACH sites x 50
reads, HCT116_RNA002 weights, num_iterations=1000, 20 reads per bag.  --config selects the other BASELINE shapes
(2: 100k x 20; 4: 500k x 30, HEK293T_RNA004; 5: 250k pooled sites x 80 reads = 4 input directories x 20, arabidopsis).
For N > 1 the SAME job is sharded by contiguous site ranges (strong scaling), one NCCL all-gather of the per-site
outputs at the end of every step; the kernel writes its (site_prob, mod_count) pairs straight into the send buffer.

A "step" is one pass of the hot path over the whole job.  One JSON line is printed by rank 0:
  value         sites/s with inputs resident in HBM (CUDA events, max over ranks)
  e2e           sites/s through the host-buffer C-ABI call (pinned host inputs, H2D + kernel + D2H pipelined inside the
                timed region)
  roofline      achieved algorithmic GB/s of the fused kernel vs the measured HBM peak (+ executed FLOP and MC draw rates)
  parity        outside the timed region: >= 50 000 sites of THIS job (first, last and strided runs of every shard)
                against the C oracle on the shared index stream; the run FAILS above 1e-4 (site) / 3e-6 (read vs float64)
  result_digest sha256 of the gathered site_prob || mod_count and of the per-read probabilities (fixed 8-block layout):
                identical for N = 1, 2, 4, 8 when sharding does not change results
  cpu_baseline  the reference's own CPU code (oracle/_ref) or, when absent, the oracle's port, on a bounded sample
`--impl reference` times the reference's CPU path on the same config and prints the same line shape.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "DRACH sites/sec at num_iterations=1000"
UNIT = "sites/s"
MODELS = {
    "HCT116_RNA002": ("rna002_hct116.npz", 0.033379376),
    "arabidopsis_RNA002": ("rna002_arabidopsis_virc.npz", 0.0032978046219796),
    "HEK293T_RNA004": ("rna004_hek293t_glori.npz", 0.033379376),
}
# BASELINE.json configs (index = position in `configs`, 1-based like SURVEY.md section 8d)
CONFIGS = {
    2: dict(sites=100_000, reads=20, model="HCT116_RNA002"),
    3: dict(sites=1_000_000, reads=50, model="HCT116_RNA002"),
    4: dict(sites=500_000, reads=30, model="HEK293T_RNA004"),
    5: dict(sites=250_000, reads=80, model="arabidopsis_RNA002"),
}
FALLBACK_HBM_GBS = 6650.0   # /opt/skills/guides/B200_PROFILING.md fallback
SITE_BAR, READ_BAR = 1e-4, 3e-6
DIGEST_BLOCKS = 8


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=sorted(CONFIGS))
    ap.add_argument("--sites", type=int, default=None)
    ap.add_argument("--reads", type=int, default=None)
    ap.add_argument("--iters", type=int, default=1000)
    ap.add_argument("--model", default=None, choices=sorted(MODELS))
    ap.add_argument("--encoder", default=None, choices=["tc", "ffma"], help="read-encoder implementation (default: the library's)")
    ap.add_argument("--cpu-sample-sites", type=int, default=20_000)
    ap.add_argument("--parity-sites", type=int, default=50_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-digest", action="store_true")
    ap.add_argument("--cli", action="store_true",
                    help="also run the drop-in CLI leg (SURVEY.md 8f rows 1-2): a synthetic data.json of --cli-sites x --reads "
                         "through `python -m m6anet_b200 inference` into both CSVs; wall time and per-stage rates under the "
                         "key `cli` (rank 0, N=1 only; tools/gpu_cli_timing.py)")
    ap.add_argument("--cli-sites", type=int, default=100_000)
    ap.add_argument("--tile-reads", type=int, default=0, help="override the kernel's feature rows per tile (0 = automatic)")
    ap.add_argument("--ragged", action="store_true", help="robustness run: lognormal n_reads (median 33, clip [20, 1000]) "
                    "instead of the constant --reads of the headline job (SURVEY.md section 8d)")
    a = ap.parse_args()
    preset = CONFIGS[a.config]
    a.sites = preset["sites"] if a.sites is None else a.sites
    a.reads = preset["reads"] if a.reads is None else a.reads
    a.model = preset["model"] if a.model is None else a.model
    return a


def workload_name(a):
    reads = "ragged reads (lognormal, median 33, 20..1000)" if a.ragged else f"{a.reads} reads"
    extra = " (4 input directories x 20 reads pooled per site)" if a.config == 5 and not a.ragged else ""
    return (f"synthetic {a.sites} DRACH sites x {reads}{extra}, {a.model}, num_iterations={a.iters}, 20 reads/bag, "
            f"site-sharded")


def config_dict(a):
    """The workload description BOTH arms print (a function of the command line only)."""
    per_rank_mb = a.sites * a.reads * 36 / max(a.gpus, 1) / 1e6
    return {"workload": workload_name(a), "baseline_config": a.config, "sites": a.sites, "reads_per_site": a.reads,
            "num_iterations": a.iters, "n_samples": 20, "weights": a.model,
            "parallelism": f"site-sharded x{a.gpus}" + (", 1 NCCL all-gather/step" if a.gpus > 1 else ""),
            "l2": (f"inputs ({per_rank_mb:.0f} MB/rank) larger than L2, no flush" if per_rank_mb > 126
                   else f"inputs ({per_rank_mb:.0f} MB/rank) fit L2 (not flushed)")}


def algorithmic_bytes_per_site(n_reads: int) -> int:
    # features in + k-mer ids + CSR offset + per-read prob out + site_prob + mod_count
    return n_reads * 36 + 12 + 8 + n_reads * 4 + 8


def _synth_block(block: int, site_a: int, site_b: int, n_reads: int, seed_tag: int, ragged: bool):
    """One generation block [site_a, site_b): N(0,1) features, uniform valid k-mer ids."""
    rng = np.random.default_rng([0, seed_tag, block])
    ns = site_b - site_a
    if ragged:
        n = np.clip(np.round(np.exp(rng.normal(np.log(33), 0.8, size=ns))), 20, 1000).astype(np.int64)
        feats = rng.standard_normal((int(n.sum()), 9), dtype=np.float32)
        return feats, n, rng.integers(0, 66, size=(ns, 3), dtype=np.int32)
    feats = rng.standard_normal((ns * n_reads, 9), dtype=np.float32)
    return feats, np.full(ns, n_reads, dtype=np.int64), rng.integers(0, 66, size=(ns, 3), dtype=np.int32)


def synth_shard(site_a: int, site_b: int, n_reads: int, seed_tag: int, ragged: bool = False, total_sites: int = 0):
    """Sites [site_a, site_b) of THE job.  The job is generated in DIGEST_BLOCKS fixed blocks of sites, each from its own
    seed, so that every sharding (N = 1, 2, 4, 8, or any other) scores bit-identical inputs; a shard generates the blocks it
    overlaps and keeps its slice."""
    total = total_sites or site_b
    edges = [total * b // DIGEST_BLOCKS for b in range(DIGEST_BLOCKS + 1)]
    fs, ns_, ks = [], [], []
    for b in range(DIGEST_BLOCKS):
        lo, hi = edges[b], edges[b + 1]
        if hi <= site_a or lo >= site_b or hi == lo:
            continue
        f, n, k = _synth_block(b, lo, hi, n_reads, seed_tag, ragged)
        a_, b_ = max(site_a, lo) - lo, min(site_b, hi) - lo
        off = np.concatenate([[0], np.cumsum(n)])
        fs.append(f[off[a_]:off[b_]])
        ns_.append(n[a_:b_])
        ks.append(k[a_:b_])
    if not fs:
        return np.zeros((0, 9), np.float32), np.zeros(1, np.int64), np.zeros((0, 3), np.int32)
    n_all = np.concatenate(ns_)
    read_off = np.concatenate([[0], np.cumsum(n_all)]).astype(np.int64)
    return np.ascontiguousarray(np.concatenate(fs)), read_off, np.ascontiguousarray(np.concatenate(ks))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons with NVML during the timed region."""

    def __init__(self, index: int, period_s: float = 0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        return {"sm_mhz": (statistics.median(self.samples) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- the CPU arm ---------------------------------------------------------------------------------------------------------
def cpu_reference_run(a, feats, off, kmer, cores):
    """One pass of the reference's CPU hot path over the given sites.  The reference's own code (oracle/_ref, installed by
    oracle/make_ref.sh) when present -- kind "reference" --, else the oracle's port of it -- kind "port"."""
    from oracle import ref_arm
    if ref_arm.available():
        r = ref_arm.time_stock_functions(a.model, feats, off, kmer, a.iters, cores)
        how = ("stock MILModel.get_read_representation + probability_layer in 16-site batches, group_results, "
               "calculate_site_proba(n_processes=%d) once" % cores)
        return "reference", r, how
    from oracle import ReadEncoderParams
    from oracle.cpu_baseline import time_reference_port
    params = ReadEncoderParams.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", MODELS[a.model][0]))
    r = time_reference_port(params, feats, off, kmer, a.iters, n_procs=cores, read_threshold=MODELS[a.model][1])
    return "port", r, "oracle port of the reference's torch / NumPy calls, one Pool over all cores"


def run_reference(a):
    """--impl reference: the reference's CPU path on all host cores, each step a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle.cpu_baseline import host_cores
    n_sample = min(a.sites, a.cpu_sample_sites)
    feats, off, kmer = synth_shard(0, n_sample, a.reads, 1, total_sites=a.sites)
    cores = host_cores()
    for _ in range(min(a.warmup, 1)):
        cpu_reference_run(a, feats, off, kmer, cores)
    times, last, kind, how = [], None, "port", ""
    for _ in range(a.steps):
        kind, last, how = cpu_reference_run(a, feats, off, kmer, cores)
        times.append(n_sample / last["sites_per_s"])
    t = sum(times)
    value = n_sample * a.steps / t
    sample = (f"{n_sample} sites x {a.reads} reads per step (bounded sample of the {a.sites}-site job): {how}; "
              f"encoder={last['t_encoder_s']:.3f}s mc={last['t_mc_s']:.3f}s")
    extra = {}
    if kind == "reference":       # context: the literal stock loop (Pool re-forked at every flush, CSV rows) on a smaller sample
        try:
            from oracle import ref_arm
            n_lit = min(n_sample, 2048)
            lit = ref_arm.time_stock_run_inference(a.model, feats[: int(off[n_lit])], off[: n_lit + 1], kmer[:n_lit], a.iters, cores)
            extra["stock_run_inference"] = {"value": lit["sites_per_s"], "unit": UNIT, "sites": lit["sites_given"],
                                            "sites_written": lit["sites_written"],
                                            "note": "the reference's run_inference() itself on resident inputs (flush cadence, "
                                                    "Pool per flush, CSV rows); its flush rule drops the last batches"}
        except Exception as exc:                                    # noqa: BLE001 - context only
            extra["stock_run_inference"] = {"error": repr(exc)}
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * t / a.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": config_dict(a),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, **extra},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- parity of the timed job against the oracle (outside the timed region) -----------------------------------------------
def parity_runs(ns: int, want_sites: int):
    """Contiguous runs of sites covering the shard: the first, the last and strided ones in between."""
    if ns <= want_sites:
        return [(0, ns)]
    run = 1000
    n_runs = max(3, -(-want_sites // run))
    starts = np.unique(np.linspace(0, ns - run, n_runs).astype(np.int64))
    return [(int(s), int(min(s + run, ns))) for s in starts]


def parity_check(a, sa, feats_h, off_h, kmer_h, rp_dev, sp_dev, mc_dev, thr, want_sites):
    from oracle import ReadEncoderParams, c_oracle, read_probabilities_float64
    params = ReadEncoderParams.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", MODELS[a.model][0]))
    worst_site = worst_read64 = worst_read32 = 0.0
    mism = n_sites = n_reads = 0
    for lo, hi in parity_runs(len(off_h) - 1, want_sites):
        r_lo, r_hi = int(off_h[lo]), int(off_h[hi])
        f, o, k = feats_h[r_lo:r_hi], off_h[lo:hi + 1] - off_h[lo], kmer_h[lo:hi]
        orp, osp, omc = c_oracle.mil_inference(params, f, o, k, a.iters, seed=0, site_id_base=sa + lo, n_samples=20,
                                               read_threshold=thr)
        rows = None if params.emb is None else np.repeat(k, np.diff(o), axis=0)
        p64 = read_probabilities_float64(params, f, rows)
        rp = rp_dev[r_lo:r_hi]
        worst_read64 = max(worst_read64, float(np.abs(rp - p64).max()))
        worst_read32 = max(worst_read32, float(np.abs(rp - orp).max()))
        worst_site = max(worst_site, float(np.abs(sp_dev[lo:hi] - osp).max()))
        # a read within 1e-6 of the threshold may be counted either way by two float32 evaluations
        near = np.abs(orp.astype(np.float64) - float(np.float32(thr))) < 1e-6
        slack = np.add.reduceat(near.astype(np.int64), o[:-1]) if len(o) > 1 else np.zeros(0, np.int64)
        mism += int(np.sum(np.abs(mc_dev[lo:hi].astype(np.int64) - omc) > slack))
        n_sites += hi - lo
        n_reads += r_hi - r_lo
    return {"sites": n_sites, "reads": n_reads, "max_abs_site": worst_site, "max_abs_read": worst_read64,
            "max_abs_read_vs_float32_oracle": worst_read32, "mod_count_mismatch": mism}


def digest_blocks(total_sites: int):
    return [total_sites * b // DIGEST_BLOCKS for b in range(DIGEST_BLOCKS + 1)]


def main():
    a = parse_args()
    if a.impl == "reference":
        return run_reference(a)

    import torch
    import torch.distributed as dist
    from m6anet_b200 import _cabi, weights as W
    from m6anet_b200.engine import MilEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    a.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_cpus = []
    if world > 1 and os.environ.get("M6A_NO_NUMA_BIND") != "1":
        # one process per GPU: stay on the GPU's socket so that pinned staging buffers are first-touched next to it
        from m6anet_b200.dist import bind_host_to_device
        numa_cpus = bind_host_to_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner with printf on fd 1 at communicator creation; stdout must carry only the JSON
        # line, so fd 1 points at stderr until the first collective has run.
        sys.stdout.flush()
        saved_fd1 = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd1, 1)
            os.close(saved_fd1)
    n_gpus = world

    # ---- shard: contiguous site ranges, equal read counts (constant n_reads) -------------------------
    bounds = [a.sites * r // world for r in range(world + 1)]
    sa, sb = bounds[rank], bounds[rank + 1]
    ns = sb - sa
    shard_max = max(bounds[r + 1] - bounds[r] for r in range(world))
    feats_h, off_h, kmer_h = synth_shard(sa, sb, a.reads, 1, a.ragged, total_sites=a.sites)
    thr = MODELS[a.model][1]
    eng = MilEngine(W.from_npz(os.path.join(ROOT, "m6anet_b200", "assets", "model_states", MODELS[a.model][0])), dev)
    if a.encoder:
        eng.set_encoder(a.encoder)
    if a.tile_reads:
        eng.set_tile_reads(a.tile_reads)

    # host buffers: page-locked through the library (the same allocator run_inference uses), not torch
    feats_p = _cabi.pinned_empty(feats_h.shape, np.float32)
    off_p = _cabi.pinned_empty(off_h.shape, np.int64)
    kmer_p = _cabi.pinned_empty(kmer_h.shape, np.int32)
    feats_p[...], off_p[...], kmer_p[...] = feats_h, off_h, kmer_h
    feats_d, off_d, kmer_d = (torch.from_numpy(x).to(dev) for x in (feats_p, off_p, kmer_p))
    read_prob = torch.empty(feats_d.shape[0], dtype=torch.float32, device=dev)
    # per-site outputs interleaved [shard_max, 2] (site_prob, mod_count bits): written by the kernel, sent by ONE all-gather
    site_pack = torch.zeros((shard_max, 2), dtype=torch.float32, device=dev)
    gathered = torch.empty((world * shard_max, 2), dtype=torch.float32, device=dev) if world > 1 else None

    def step():
        eng.infer_device_packed(feats_d, off_d, kmer_d, a.iters, seed=0, site_id_base=sa, read_threshold=thr,
                                read_prob=read_prob, site_out=site_pack)
        if world > 1:
            dist.all_gather_into_tensor(gathered, site_pack)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(a.warmup, 0)):
        step()
    barrier()

    sampler = ClockSampler(local_rank)
    sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.steps)]
    barrier()
    ev[0].record()
    for i in range(a.steps):
        kev[i][0].record()
        eng.infer_device_packed(feats_d, off_d, kmer_d, a.iters, seed=0, site_id_base=sa, read_threshold=thr,
                                read_prob=read_prob, site_out=site_pack)
        kev[i][1].record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, site_pack)
    ev[1].record()
    barrier()
    clocks = sampler.stop()
    elapsed_ms = ev[0].elapsed_time(ev[1])
    kernel_ms = [k0.elapsed_time(k1) for k0, k1 in kev]
    launch = eng.last_launch()
    launch["encoder"] = eng.encoder
    per_rank = None
    if world > 1:
        t = torch.tensor([elapsed_ms, max(kernel_ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t[0])
        mine = torch.tensor([sum(kernel_ms) / len(kernel_ms), float(clocks["sm_mhz"] or 0), float(len(clocks["reasons"]))],
                            dtype=torch.float64, device=dev)
        allr = torch.empty((world, 3), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, mine)
        per_rank = {"kernel_ms": [round(float(v), 3) for v in allr[:, 0]], "sm_mhz": [float(v) for v in allr[:, 1]],
                    "n_throttle_reasons": [int(v) for v in allr[:, 2]]}
    ms_per_step = elapsed_ms / a.steps
    value = a.sites / (ms_per_step * 1e-3)

    # ---- roofline of the fused kernel (this rank's launches; algorithmic bytes / CUDA-event time) -----
    peak, peak_src = peaks()
    k_avg_ms = sum(kernel_ms) / len(kernel_ms)
    n_reads_rank = int(feats_h.shape[0])
    alg_bytes = n_reads_rank * 40 + ns * 28 if a.ragged else ns * algorithmic_bytes_per_site(a.reads)
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1:
        try:
            tj = json.load(open(tpath))
            if tj.get("sites") == a.sites and tj.get("reads") == a.reads and not a.ragged and a.iters == tj.get("iters"):
                tk = tj if tj.get("encoder") == eng.encoder else tj.get(eng.encoder) or {}
                traffic = tk.get("dram_bytes_per_launch")
                traffic_note = tk.get("source")
        except Exception:
            pass
    kernel_name = "mil_infer_tc_kernel" if eng.encoder == "tc" else "mil_infer_kernel"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_note, "kernel": kernel_name, "kernel_ms": k_avg_ms,
                "step_ms": ms_per_step, "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "per_rank": per_rank,
                "note": "not HBM-bound (DRAM ~2.4 % busy): the Monte-Carlo pooling (20 k draws per site) keeps the FMA-heavy "
                        "pipe ~60 % busy with three quarter-rate integer multiplies per pair of draws, the encoder roles share "
                        "the same SM; see DESIGN.md section 4 and profiles/r02_final_tc_ncu_raw_summary.txt"}
    # Executed arithmetic of one pass (SURVEY 8d lists the reference formulation, 14 164 FLOP/read; the kernel executes
    # 2 x (150*9 + 150*32 + 32) = 12 364 FLOP per read + 900 MAC per site of embedding fold) and the Monte-Carlo draw rate.
    try:
        props = torch.cuda.get_device_properties(dev)
        sm_mhz = float(clocks["sm_max_mhz"] or 1965)
        fp32_peak = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
        flop_exec = n_reads_rank * 12364.0 + ns * 1800.0
        roofline["compute"] = {
            "executed_flop_per_read": 12364, "executed_flop_per_site_extra": 1800, "samples_per_site": 20 * a.iters,
            "achieved_tflops_fp32_equivalent": flop_exec / (k_avg_ms * 1e-3) / 1e12,
            "peak_tflops_fp32_cuda_cores": fp32_peak,
            "frac_of_cuda_core_peak": flop_exec / (k_avg_ms * 1e-3) / 1e12 / fp32_peak,
            "mc_draws_per_s": ns * 20.0 * a.iters / (k_avg_ms * 1e-3),
            "where": ("Linear blocks on tcgen05 tensor cores as 3xTF32 (3 MMA terms per product: 3x the FLOP above on the "
                      "tensor pipe); CUDA cores run relu/split, sigmoid and the Monte-Carlo pooling" if eng.encoder == "tc"
                      else "Linear blocks on CUDA cores (FFMA2)"),
            "peak_source": f"{props.multi_processor_count} SMs x 128 FP32 lanes x 2 x {sm_mhz:.0f} MHz (nominal)"}
    except Exception as exc:      # informative only: never fail the bench line over it
        roofline["compute"] = {"error": repr(exc)}

    # ---- results of the timed job: digest + parity against the oracle (outside the timed region) ---------------------------
    step()
    barrier()
    pack_h = site_pack[:ns].cpu().numpy()
    sp_h = np.ascontiguousarray(pack_h[:, 0])
    mc_h = np.ascontiguousarray(pack_h[:, 1]).view(np.int32)
    rp_h = read_prob.cpu().numpy()
    digest = None
    if not a.no_digest:
        blocks = digest_blocks(a.sites)
        aligned = all(b in blocks for b in (sa, sb)) and not a.ragged
        if world > 1:
            g = gathered.view(world, shard_max, 2).cpu().numpy()
            sp_all = np.concatenate([g[r, : bounds[r + 1] - bounds[r], 0] for r in range(world)])
            mc_all = np.concatenate([g[r, : bounds[r + 1] - bounds[r], 1] for r in range(world)]).view(np.int32)
        else:
            sp_all, mc_all = sp_h, mc_h
        site_digest = hashlib.sha256(np.ascontiguousarray(sp_all).tobytes() + np.ascontiguousarray(mc_all).tobytes()).hexdigest()
        read_digest = None
        if aligned:      # per-read probabilities: one sha256 per fixed block of sites, concatenated in block order
            mine = []
            for b in range(DIGEST_BLOCKS):
                lo, hi = blocks[b], blocks[b + 1]
                if lo >= sa and hi <= sb:
                    mine.append((b, hashlib.sha256(rp_h[int(off_h[lo - sa]):int(off_h[hi - sa])].tobytes()).digest()))
            if world > 1:
                buf = torch.zeros((DIGEST_BLOCKS, 32), dtype=torch.uint8, device=dev)
                for b, d in mine:
                    buf[b] = torch.frombuffer(bytearray(d), dtype=torch.uint8).to(dev)
                dist.all_reduce(buf, op=dist.ReduceOp.SUM)
                cat = buf.cpu().numpy().tobytes()
            else:
                cat = b"".join(d for _, d in mine)
            read_digest = hashlib.sha256(cat).hexdigest()
        digest = {"site_prob_mod_count_sha256": site_digest, "read_prob_sha256": read_digest,
                  "layout": f"site arrays in global site order; read_prob as {DIGEST_BLOCKS} per-block sha256 digests of equal "
                            f"site ranges, hashed in order (comparable for N in 1, 2, 4, 8)"}

    parity = None
    if not a.no_parity:
        want = max(1000, a.parity_sites // world)
        parity = parity_check(a, sa, feats_h, off_h, kmer_h, rp_h, sp_h, mc_h, thr, want)
        if world > 1:
            t = torch.tensor([parity["max_abs_site"], parity["max_abs_read"], parity["max_abs_read_vs_float32_oracle"]],
                             dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            c = torch.tensor([parity["sites"], parity["reads"], parity["mod_count_mismatch"]], dtype=torch.int64, device=dev)
            dist.all_reduce(c, op=dist.ReduceOp.SUM)
            parity = {"sites": int(c[0]), "reads": int(c[1]), "max_abs_site": float(t[0]), "max_abs_read": float(t[1]),
                      "max_abs_read_vs_float32_oracle": float(t[2]), "mod_count_mismatch": int(c[2])}
        parity["bars"] = {"site": SITE_BAR, "read_vs_float64": READ_BAR}
        parity["checker"] = "oracle/c on the shared index stream; first, last and strided 1000-site runs of every shard"
        parity["ok"] = bool(parity["max_abs_site"] <= SITE_BAR and parity["max_abs_read"] <= READ_BAR
                            and parity["mod_count_mismatch"] == 0)

    # ---- e2e: host buffers through the C-ABI host call, H2D + kernel + D2H inside the timed region ------
    e2e = None
    if not a.no_e2e:
        rp_o = _cabi.pinned_empty(feats_h.shape[0], np.float32)
        sp_o = _cabi.pinned_empty(ns, np.float32)
        mc_o = _cabi.pinned_empty(ns, np.int32)
        e_steps = max(1, min(a.steps, 5))
        eng.infer_host(feats_p, off_p, kmer_p, a.iters, seed=0, site_id_base=sa, read_threshold=thr, out=(rp_o, sp_o, mc_o))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            eng.infer_host(feats_p, off_p, kmer_p, a.iters, seed=0, site_id_base=sa, read_threshold=thr, out=(rp_o, sp_o, mc_o))
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t[0])
        same = bool(np.array_equal(sp_o, sp_h) and np.array_equal(mc_o, mc_h) and np.array_equal(rp_o, rp_h))
        e2e = {"value": a.sites * e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(feats_h.nbytes + off_h.nbytes + kmer_h.nbytes),
               "d2h_bytes_per_step": int(rp_o.nbytes + sp_o.nbytes + mc_o.nbytes),
               "steps": e_steps, "ms_per_step": 1e3 * dt / e_steps, "bytes_are": "per rank",
               "buffers": "page-locked host memory from m6a_pinned_alloc (the allocator the CLI's batches use)",
               "timer": "host perf_counter around the synchronous C-ABI call, max over ranks",
               "matches_resident_path": same}
        e2e["h2d_gbs"] = e2e["h2d_bytes_per_step"] / (dt / e_steps) / 1e9
        ceil = h2d_ceiling(world)
        if ceil:
            e2e.update(ceil)

    # ---- CPU baseline (rank 0, N=1 only) -----------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle.cpu_baseline import host_cores
        n_sample = min(a.sites, a.cpu_sample_sites)
        kind, r, how = cpu_reference_run(a, feats_h[: int(off_h[n_sample])], off_h[: n_sample + 1], kmer_h[:n_sample],
                                         host_cores())
        cpu = {"value": r["sites_per_s"], "unit": UNIT, "cores": r["cores"], "kind": kind,
               "sample": f"first {n_sample} sites of the job ({int(off_h[n_sample])} reads): {how}; encoder "
                         f"{r['t_encoder_s']:.3f}s + MC {r['t_mc_s']:.3f}s"}

    # ---- optional CLI leg: data.json -> CSVs through the drop-in command (rank 0, N=1 only) ----------------------------
    cli = None
    if rank == 0 and world == 1 and a.cli:
        import subprocess
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gpu_cli_timing.py"), str(a.cli_sites), str(a.reads)],
                           capture_output=True, text=True, cwd=ROOT)
        try:
            cli = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception:                                            # noqa: BLE001
            cli = {"error": (r.stderr or r.stdout)[-400:]}

    rc = 0
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": config_dict(a),
            "run": {"launch": launch, "build": _cabi.lib().m6a_build_info().decode(),
                    "host_binding": (f"rank 0 bound to {len(numa_cpus)} GPU-local CPUs (NVML affinity)" if numa_cpus
                                     else "none")},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "result_digest": digest,
            "clocks": clocks, "gpu_launches": a.steps * launch["n_launches"],
        }
        if cli is not None:
            line["cli"] = cli
        print(json.dumps(line), flush=True)
        if parity is not None and not parity["ok"]:
            print(f"bench.py: PARITY FAILED {parity}", file=sys.stderr)
            rc = 1
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return rc


def h2d_ceiling(world: int):
    """Concurrent pinned H2D ceiling of this box for `world` GPUs (tools/microbench/h2d_concurrent.cu, measured once and
    committed under profiles/): lets the e2e number carry its own roofline fraction."""
    p = os.path.join(ROOT, "profiles", "r02_h2d_ceiling.json")
    try:
        tj = json.load(open(p))
        row = tj["concurrent_h2d_gbs"].get(str(world))
        if row:
            return {"h2d_ceiling_gbs_per_gpu": row["per_gpu"], "h2d_ceiling_gbs_aggregate": row["aggregate"],
                    "h2d_ceiling_source": "profiles/r02_h2d_ceiling.json"}
    except Exception:
        pass
    return None


if __name__ == "__main__":
    sys.exit(main())
