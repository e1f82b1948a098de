"""`m6anet inference` on the B200 kernel: argparser(), main(args), run_inference(model, dl, args).

Drop-in mirror of reference m6anet/scripts/inference.py (argparser :20-67, main :70-106) and
m6anet/utils/inference_utils.py (run_inference :14-71, group_results :107-140): same flags, same
pretrained-model resolution, same inputs (data.info/data.json) and the same two outputs
    data.site_proba.csv   transcript_id,transcript_position,n_reads,probability_modified,kmer,mod_ratio
    data.indiv_proba.csv  transcript_id,transcript_position,read_index,probability_modified
with the reference's row formats ('%s,%d,%s,%.16f,%s,%.16f' / '%s,%d,%s,%.16f', inference_utils.py:59-67).

Differences, all deliberate (SURVEY.md section 3.1 "do not reproduce"):
  * every site is written (the reference's `if (it + 1) % save_per_batch` flush drops trailing batches);
  * the Monte-Carlo stream is the counter-based device stream keyed by (--seed, global site index), so two
    runs -- on any number of GPUs -- give identical files (the reference forks an unseeded Pool per flush);
  * single-directory read_index is written as an integer (the reference prints the float "966210.0");
  * --device defaults to "cuda": the hot path has no CPU implementation here and refuses "cpu".
Under torchrun (WORLD_SIZE > 1) sites are sharded by read count over the ranks, each rank scores its
shard, one all-gather collects the per-site outputs and the shard files are concatenated in site order.
"""
from __future__ import annotations

import os
import pathlib
import shutil
import warnings
from argparse import ArgumentDefaultsHelpFormatter, ArgumentParser
from typing import List

import numpy as np

from .constants import (DEFAULT_MIN_READS, DEFAULT_MODEL_CONFIG, DEFAULT_NORM_PATH, DEFAULT_PRETRAINED_MODEL,
                        DEFAULT_PRETRAINED_MODELS, DEFAULT_READ_THRESHOLD, PRETRAINED_CONFIGS)
from .data import NanopolishDS, NanopolishReplicateDS, SiteBatch
from .dist import all_gather_site_outputs, env_world, shard_bounds
from .model import MILModel, load_model_config

SITE_HEADER = 'transcript_id,transcript_position,n_reads,probability_modified,kmer,mod_ratio\n'
INDIV_HEADER = 'transcript_id,transcript_position,read_index,probability_modified\n'
N_SAMPLES = 20                       # the literal 20 of reference utils/inference_utils.py:54
DEFAULT_READS_PER_BATCH = 2_000_000  # ingest/score/write granularity (72 MB of features)


def argparser():
    parser = ArgumentParser(formatter_class=ArgumentDefaultsHelpFormatter, add_help=False)
    # Required arguments
    parser.add_argument("--input_dir", nargs="*", help='directories containing data.info and data.json.', required=True)
    parser.add_argument("--out_dir", help='directory to output inference results.', required=True)
    # Optional arguments (reference scripts/inference.py:33-66)
    parser.add_argument("--pretrained_model",
                        help="pre-trained model available at m6anet. Options include {}.".format(DEFAULT_PRETRAINED_MODELS),
                        default=DEFAULT_PRETRAINED_MODEL, type=str)
    parser.add_argument("--model_config", help='path to model config file.', default=DEFAULT_MODEL_CONFIG)
    parser.add_argument("--model_state_dict", help='path to model weights.', default=None)
    parser.add_argument("--norm_path", help='path to normalization factors file', default=DEFAULT_NORM_PATH)
    parser.add_argument("--batch_size", help='batch size for inference (sites per reference batch; kept for '
                        'compatibility, the kernel batches by --reads_per_batch).', default=16, type=int)
    parser.add_argument("--save_per_batch", help='saving inference results every save_per_batch multiples '
                        '(kept for compatibility; every site is always written).', default=2, type=int)
    parser.add_argument("--n_processes", help='number of processes to run (here: ingest / CSV worker threads).', default=25,
                        type=int)
    parser.add_argument("--num_iterations", help='number of sampling run.', default=1000, type=int)
    parser.add_argument("--device", help='device to perform inference with (cuda or cuda:N).', default='cuda', type=str)
    parser.add_argument("--seed", help='random seed for sampling.', default=0, type=int)
    parser.add_argument("--read_proba_threshold",
                        help='default probability threshold for a read to be considered modified.',
                        default=DEFAULT_READ_THRESHOLD, type=float)
    # m6anet_b200 extension
    parser.add_argument("--reads_per_batch", help='reads ingested, scored and written per step.',
                        default=DEFAULT_READS_PER_BATCH, type=int)
    return parser


# ---- CSV emit (native, multi-threaded: m6a_write_site_csv / m6a_write_indiv_csv) -------------------------
def _tx_buffer(batch: SiteBatch):
    """(concatenated utf-8 transcript ids, CSR offsets [S+1]) of a batch, vectorised and cached on the batch
    (both writers need it; a per-site Python loop costs ~2 us per site)."""
    cached = getattr(batch, "_tx_buffer", None)
    if cached is not None:
        return cached
    ids = np.asarray(batch.tx_ids if batch.tx_bytes is None else batch.tx_bytes)
    if ids.dtype.kind != "S":
        ids = np.char.encode(ids.astype(str), "utf-8") if len(ids) else np.zeros(0, dtype="S1")
    width = max(ids.dtype.itemsize, 1)
    lengths = np.char.str_len(ids).astype(np.int64) if len(ids) else np.zeros(0, dtype=np.int64)
    mat = np.ascontiguousarray(ids).view(np.uint8).reshape(len(ids), width) if len(ids) else np.zeros((0, 1), np.uint8)
    buf = mat[np.arange(width)[None, :] < lengths[:, None]].tobytes()        # row-major => ids in site order
    off = np.zeros(len(ids) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    batch._tx_buffer = (buf, off)
    return batch._tx_buffer


def _vp(a):
    import ctypes as C
    return a.ctypes.data_as(C.c_void_p)


def write_site_rows(f, batch: SiteBatch, site_prob: np.ndarray, mod_count: np.ndarray, n_threads: int = 0):
    """'%s,%d,%s,%.16f,%s,%.16f' rows (tx_id, tx_pos, n_reads, site_prob, kmer, mod_ratio) appended to file object f;
    mod_ratio = mod_count / n_reads in float64 == np.mean(p >= thr) (reference utils/inference_utils.py:53,59-60)."""
    from . import _cabi
    f.flush()
    tx, off = _tx_buffer(batch)
    kmer5 = np.asarray(batch.kmers).astype("S5").tobytes() if batch.n_sites else b""     # [S, 5] chars, NUL padded
    pos = np.ascontiguousarray(batch.tx_pos, dtype=np.int64)
    ro = np.ascontiguousarray(batch.read_off, dtype=np.int64)
    sp = np.ascontiguousarray(site_prob, dtype=np.float32)
    mc = np.ascontiguousarray(mod_count, dtype=np.int32)
    _cabi.check(_cabi.lib().m6a_write_site_csv(f.fileno(), batch.n_sites, tx, _vp(off), _vp(pos), _vp(ro), _vp(sp), _vp(mc),
                                               kmer5, n_threads), "m6a_write_site_csv")


def write_indiv_rows(g, batch: SiteBatch, read_prob: np.ndarray, n_threads: int = 0):
    """'%s,%d,%s,%.16f' rows (tx_id, tx_pos, read_index, read_prob) (reference utils/inference_utils.py:63-64);
    read_index is the integer id, or "{id}_{replicate}" for multi-directory input (utils/data_utils.py:421-423)."""
    from . import _cabi
    g.flush()
    tx, off = _tx_buffer(batch)
    pos = np.ascontiguousarray(batch.tx_pos, dtype=np.int64)
    ro = np.ascontiguousarray(batch.read_off, dtype=np.int64)
    ids = np.ascontiguousarray(batch.read_ids, dtype=np.int64)
    rep = None if batch.read_rep is None else np.ascontiguousarray(batch.read_rep, dtype=np.int32)
    rp = np.ascontiguousarray(read_prob, dtype=np.float32)
    _cabi.check(_cabi.lib().m6a_write_indiv_csv(g.fileno(), batch.n_sites, tx, _vp(off), _vp(pos), _vp(ro), _vp(ids),
                                                None if rep is None else _vp(rep), _vp(rp), n_threads), "m6a_write_indiv_csv")


def plan_batches(n_reads: np.ndarray, lo: int, hi: int, reads_per_batch: int) -> List[tuple]:
    """Cut sites [lo, hi) into consecutive spans of about reads_per_batch reads."""
    spans, a, acc = [], lo, 0
    for s in range(lo, hi):
        acc += int(n_reads[s])
        if acc >= reads_per_batch:
            spans.append((a, s + 1))
            a, acc = s + 1, 0
    if a < hi:
        spans.append((a, hi))
    return spans


def append_file(out, path: str) -> None:
    """Append the file at `path` at the current position of the open binary file `out`: in-kernel copy (sendfile) where
    the platform allows it -- the shard files of a multi-GPU run are gigabytes of CSV --, a buffered copy otherwise
    (sendfile refuses O_APPEND descriptors: open `out` with 'r+b' and seek to the end to get the fast path)."""
    out.flush()
    with open(path, 'rb') as part:
        size = os.fstat(part.fileno()).st_size
        sent = 0
        try:
            while sent < size:
                n = os.sendfile(out.fileno(), part.fileno(), sent, min(size - sent, 1 << 30))
                if n == 0:
                    break
                sent += n
        except (OSError, AttributeError):     # sendfile unsupported for this pair of descriptors
            pass
        if sent < size:
            part.seek(sent)
            shutil.copyfileobj(part, out, 1 << 24)
        elif sent:
            out.seek(0, os.SEEK_END)       # keep the Python-level position in step with the descriptor


def _resolve_device(device: str, local_rank: int, world: int) -> int:
    """--device -> CUDA device index (no torch import: the single-GPU path only needs the C ABI)."""
    from . import _cabi
    if str(device).startswith("cpu"):
        raise RuntimeError("m6anet_b200 runs the inference hot path only on CUDA devices (sm_100a); there is no CPU path. "
                           "Use --device cuda, or the reference implementation for CPU inference.")
    n = _cabi.device_count()
    if n < 1:
        raise RuntimeError("no CUDA device is visible; m6anet_b200 has no CPU fallback")
    text = str(device)
    idx = _cabi.parse_device(device) if ":" in text or not isinstance(device, str) else (local_rank if world > 1 else 0)
    if idx >= n:
        raise RuntimeError(f"cuda:{idx} is not visible ({n} device(s))")
    return idx


def run_inference(model: MILModel, dl, args):
    """Score every site of the dataset and append the two CSV files (reference utils/inference_utils.py:14-71).

    `dl` is the dataset itself or anything with a `.dataset` attribute (a DataLoader in the reference).
    Uses args.{out_dir, device, read_proba_threshold, num_iterations, n_processes, seed} like the reference.
    """
    ds = getattr(dl, "dataset", dl)
    if getattr(ds, "num_neighboring_features", 1) != getattr(model, "num_neighboring_features", 1):
        raise ValueError("run_inference: the dataset was built with num_neighboring_features = %r but the model expects %r"
                         % (getattr(ds, "num_neighboring_features", 1), getattr(model, "num_neighboring_features", 1)))
    rank, world, local_rank = env_world()
    dev = _resolve_device(args.device, local_rank, world)
    # M6A_DIST_BACKEND=gloo: the one all-gather runs on host tensors, so several ranks may share one GPU (tests on a
    # single-GPU box); default NCCL over NVLink
    backend = os.environ.get("M6A_DIST_BACKEND", "nccl")
    if world > 1:     # torch only for the multi-GPU plumbing (torch.distributed)
        import torch
        import torch.distributed as dist
        if os.environ.get("M6A_NO_NUMA_BIND") != "1":
            from .dist import bind_host_to_device
            bind_host_to_device(dev)      # ingest / CSV threads and pinned staging stay on the GPU's socket (best effort)
        torch.cuda.set_device(dev)
        if not dist.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            if backend == "nccl":
                dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
            else:
                dist.init_process_group(backend)
    seed = int(getattr(args, "seed", 0))
    n_iters = int(args.num_iterations)
    thr = float(args.read_proba_threshold)
    reads_per_batch = int(getattr(args, "reads_per_batch", DEFAULT_READS_PER_BATCH))

    bounds = shard_bounds(ds.n_reads, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    spans = plan_batches(ds.n_reads, lo, hi, reads_per_batch)
    n_threads = max(1, min(int(getattr(args, "n_processes", 1)), os.cpu_count() or 1))   # ingest / CSV worker threads

    site_path = os.path.join(args.out_dir, "data.site_proba.csv")
    indiv_path = os.path.join(args.out_dir, "data.indiv_proba.csv")
    suffix = f".rank{rank}" if world > 1 else ""
    all_sp, all_mc = [], []
    eng = model.engine(dev)
    from ._cabi import PinnedPool
    pool = PinnedPool()       # page-locked feature / output buffers, recycled batch to batch (H2D / D2H overlap the kernel)
    if world > 1:             # a direct call may find shard files of an earlier, failed run: they are opened for append
        for path in (site_path, indiv_path):
            if os.path.exists(path + suffix):
                os.remove(path + suffix)

    # Three overlapped stages (the native calls release the GIL): ingest of batch i+1 | H2D/kernel/D2H of batch i |
    # CSV emit of batch i-1.  Queues are bounded so at most ~3 batches are resident on the host.
    import queue
    import threading
    q_in: "queue.Queue" = queue.Queue(maxsize=2)
    q_out: "queue.Queue" = queue.Queue(maxsize=2)
    errors: list = []

    def ingest():
        try:
            for a, b in spans:
                if errors:
                    break
                q_in.put((a, ds.load_sites(a, b, n_threads=n_threads, alloc=pool.empty)))   # data.json -> flat buffers (native)
        except BaseException as e:   # noqa: BLE001 - re-raised in the main thread
            errors.append(e)
        finally:
            q_in.put(None)

    def emit():
        try:
            with open(site_path + suffix, 'ab') as f, open(indiv_path + suffix, 'ab') as g:
                while True:
                    item = q_out.get()
                    if item is None:
                        break
                    batch, read_prob, site_prob, mod_count = item
                    write_site_rows(f, batch, site_prob, mod_count, n_threads)
                    write_indiv_rows(g, batch, read_prob, n_threads)
                    pool.give_back(batch.feats, read_prob)
        except BaseException as e:   # noqa: BLE001
            errors.append(e)
            while q_out.get() is not None:    # keep draining so the main thread never blocks on a full queue
                pass

    t_in = threading.Thread(target=ingest, name="m6a-ingest", daemon=True)
    t_out = threading.Thread(target=emit, name="m6a-emit", daemon=True)
    t_in.start()
    t_out.start()
    try:
        while True:
            item = q_in.get()
            if item is None:
                break
            a, batch = item
            if errors:
                continue
            read_prob = pool.empty(batch.feats.shape[0], np.float32)
            site_prob = np.empty(batch.n_sites, np.float32)
            mod_count = np.empty(batch.n_sites, np.int32)
            eng.infer_host(batch.feats, batch.read_off, batch.kmer_idx, n_iters, seed=seed, site_id_base=a,
                           n_samples=N_SAMPLES, read_threshold=thr, out=(read_prob, site_prob, mod_count))   # H2D -> kernel -> D2H
            all_sp.append(site_prob)
            all_mc.append(mod_count)
            q_out.put((batch, read_prob, site_prob, mod_count))
    except BaseException as e:   # noqa: BLE001 - the GPU stage failed: stop the producer, then re-raise below
        errors.insert(0, e)
        while q_in.get() is not None:     # unblock an ingest thread waiting on the bounded queue
            pass
    finally:
        q_out.put(None)
        t_in.join()
        t_out.join()
    if world > 1:
        # every rank learns whether any rank failed BEFORE the collective: a rank that raised alone would leave the others
        # blocked in the all-gather until the watchdog fires
        import torch.distributed as dist
        tdev = torch.device("cuda", dev) if backend == "nccl" else torch.device("cpu")
        failed = torch.tensor([1 if errors else 0], dtype=torch.int32, device=tdev)
        dist.all_reduce(failed, op=dist.ReduceOp.MAX)
        if int(failed.item()) and not errors:
            errors.append(RuntimeError("run_inference: another rank failed; see its traceback"))
        if errors:
            for path in (site_path, indiv_path):
                if os.path.exists(path + suffix):
                    os.remove(path + suffix)
    if errors:
        raise errors[0]
    site_prob = np.concatenate(all_sp) if all_sp else np.zeros(0, np.float32)
    mod_count = np.concatenate(all_mc) if all_mc else np.zeros(0, np.int32)

    if world > 1:
        sp_t, mc_t = all_gather_site_outputs(torch.from_numpy(site_prob).to(tdev), torch.from_numpy(mod_count).to(tdev), bounds)
        site_prob, mod_count = sp_t.cpu().numpy(), mc_t.cpu().numpy()
        dist.barrier()
        if rank == 0:    # concatenate the shard files in site order behind the headers
            for path in (site_path, indiv_path):
                with open(path, 'r+b') as out:       # not 'ab': sendfile refuses O_APPEND descriptors
                    out.seek(0, os.SEEK_END)
                    for r in range(world):
                        append_file(out, f"{path}.rank{r}")
                        os.remove(f"{path}.rank{r}")
        dist.barrier()
    return site_prob, mod_count


def main(args):
    input_dir = args.input_dir
    if args.model_state_dict is not None:
        warnings.warn("--model_state_dict is specified, overwriting default model weights")
    else:
        if args.pretrained_model not in DEFAULT_PRETRAINED_MODELS:
            raise ValueError("Invalid pretrained model {}, must be one of {}".format(args.pretrained_model,
                                                                                     DEFAULT_PRETRAINED_MODELS))
        args.model_state_dict = PRETRAINED_CONFIGS[args.pretrained_model][0]
        args.read_proba_threshold = PRETRAINED_CONFIGS[args.pretrained_model][1]
        args.norm_path = PRETRAINED_CONFIGS[args.pretrained_model][2]

    model = MILModel(load_model_config(args.model_config)).to(args.device)
    model.load_weights(args.model_state_dict)

    rank, world, _ = env_world()
    pathlib.Path(args.out_dir).mkdir(parents=True, exist_ok=True)
    if rank == 0:      # headers truncate the outputs (reference scripts/inference.py:94-97)
        with open(os.path.join(args.out_dir, "data.site_proba.csv"), 'w', encoding='utf-8') as f:
            f.write(SITE_HEADER)
        with open(os.path.join(args.out_dir, "data.indiv_proba.csv"), 'w', encoding='utf-8') as g:
            g.write(INDIV_HEADER)
    for name in ("data.site_proba.csv", "data.indiv_proba.csv"):
        stale = os.path.join(args.out_dir, f"{name}.rank{rank}")
        if world > 1 and os.path.exists(stale):
            os.remove(stale)

    if len(input_dir) == 1:
        ds = NanopolishDS(input_dir[0], DEFAULT_MIN_READS, args.norm_path, mode='Inference')
    else:
        ds = NanopolishReplicateDS(input_dir, DEFAULT_MIN_READS, args.norm_path, mode='Inference')
    run_inference(model, ds, args)
