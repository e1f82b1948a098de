#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Only runs in the build container (needs /root/reference); the fixtures it writes are committed
so that the CPU suite, the GPU suite and bench.py never read /root/reference at run time.

    python tests/golden/make_golden.py

Writes
  bundled/data.json.gz, data.info, data.indiv_proba.csv.gz, data.site_proba.csv.gz
        verbatim copies of the reference's own test inputs and golden outputs
        (reference m6anet/tests/data/, used by m6anet/tests/test_inference.py:10-37).
  bundled_flat.npz
        the 101 bundled sites exactly as the reference's NanopolishDS + inference_collate
        present them to the model (utils/data_utils.py:192-231,498-506), plus the reference
        model's per-read probabilities (utils/inference_utils.py:35-37) and, for the first
        sites, `_calculate_site_proba` (utils/inference_utils.py:74-87) evaluated right after
        np.random.seed(0) (the MT19937 replay vectors).
  synthetic_inputs.npz
        288 sites (every valid 7-mer => all 66 five-mer ids), ragged n_reads in [20, 700],
        N(0,1) features with a heavy-tailed 1%.
  synthetic_<model>.npz   for the 4 registry models + a signal-only random-weight model
        per-read p from the reference model; per-site probability from the LITERAL MIL forward
        (MILModel.forward on the gathered [n_iters,20,.] bags, model/model.py:155-164, mean over
        iterations as in utils/training_utils.py:236-256) on the Philox index stream of
        oracle/philox.py; mod_count at the registry threshold.
  bundled/data.info.labelled, validate_golden.npz
        the reference's own labelled index, and the outputs of the reference's `validate()`
        (utils/training_utils.py:213-268) run UNMODIFIED on its 'Val' and 'Test' sets (NanopolishDS mode='Val'/'Test',
        DataLoader + train_collate, binary_cross_entropy_loss, 5 passes after np.random.seed) for the prod / mean / max
        pooling blocks: per-pass y_pred, y_true, roc_auc, pr_auc, avg_loss, and the bags the legacy MT19937 stream
        drew (replayed with oracle.philox.bag_indices_mt19937) so the kernel's explicit-index mode can be held
        to the very same numbers.
"""
import gzip
import json
import os
import shutil
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.modules.setdefault("ujson", json)          # the only missing dependency of `import m6anet`
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

import toml  # noqa: E402
import torch  # noqa: E402
from m6anet.model.model import MILModel  # noqa: E402
from m6anet.utils.constants import (DEFAULT_MODEL_CONFIG, KMER_TO_INT, PRETRAINED_CONFIGS)  # noqa: E402
from m6anet.utils.data_utils import NanopolishDS, inference_collate  # noqa: E402
from m6anet.utils.inference_utils import _calculate_site_proba  # noqa: E402

from oracle.philox import sample_indices  # noqa: E402

N_ITERS_GOLDEN = 200
SEED_GOLDEN = 1234
SITE_ID_BASE = 7_000_000_000      # > 2**32 so the high counter word is exercised


def ref_model(name=None, config=DEFAULT_MODEL_CONFIG, state_dict=None):
    model = MILModel(toml.load(config))
    if name is not None:
        model.load_state_dict(torch.load(PRETRAINED_CONFIGS[name][0], map_location="cpu"))
    elif state_dict is not None:
        model.load_state_dict(state_dict)
    model.eval()
    return model


def ref_read_probs(model, feats, kmer_rows):
    with torch.no_grad():
        rep = model.get_read_representation({"X": torch.from_numpy(feats), "kmer": torch.from_numpy(kmer_rows)})
        return model.pooling_filter.probability_layer(rep).flatten().numpy()


def ref_literal_site_prob(model, feats_site, kmer_site, idx):
    """MILModel.forward on [n_iters, 20, .] gathered bags, then the mean over iterations."""
    with torch.no_grad():
        X = torch.from_numpy(feats_site[idx])                       # [n_iters, 20, 9]
        K = torch.from_numpy(np.broadcast_to(kmer_site, idx.shape + (3,)).copy())
        out = model({"X": X, "kmer": K})
        return out.mean().numpy().astype(np.float32)


def bundled():
    src = os.path.join(REF, "m6anet", "tests", "data")
    dst = os.path.join(HERE, "bundled")
    os.makedirs(dst, exist_ok=True)
    with open(os.path.join(src, "data.json"), "rb") as f, gzip.GzipFile(os.path.join(dst, "data.json.gz"), "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)
    for fn in ("data.info", "data.indiv_proba.csv.gz", "data.site_proba.csv.gz"):
        shutil.copyfile(os.path.join(src, fn), os.path.join(dst, fn))

    name = "HCT116_RNA002"
    ds = NanopolishDS(src, 20, PRETRAINED_CONFIGS[name][2], mode="Inference")
    items = [ds[i] for i in range(len(ds))]
    feats, kmers, n_reads, tx_ids, tx_pos, read_ids = inference_collate(items)
    feats = feats.numpy()
    kmers = kmers.numpy()
    n_reads = n_reads.numpy()
    read_off = np.concatenate([[0], np.cumsum(n_reads)]).astype(np.int64)
    model = ref_model(name)
    p = ref_read_probs(model, feats, kmers)
    # MT19937 replay: the reference function itself, first call after np.random.seed(0)
    replay_sites = list(range(8))
    replay = []
    for s in replay_sites:
        np.random.seed(0)
        replay.append(_calculate_site_proba((p[read_off[s]:read_off[s + 1]], 1000, 20)))
    np.savez(os.path.join(HERE, "bundled_flat.npz"),
             feats=feats, kmer_idx=kmers[read_off[:-1]].astype(np.int32), read_off=read_off,
             tx_id=np.array([it[2][0] for it in items]), tx_pos=np.array([it[3][0] for it in items], dtype=np.int64),
             read_id=np.asarray(read_ids, dtype=np.float64), read_prob=p,
             replay_sites=np.array(replay_sites), replay_site_prob=np.array(replay, dtype=np.float32),
             replay_seed=0, replay_n_iters=1000)
    print("bundled:", len(ds), "sites", feats.shape[0], "reads")


def make_replicate_dirs(src_dir, out_root):
    """Two input directories for the replicate join: `a` = the bundled data, `b` = the same data.json with three
    sites missing from data.info and the last ten rows moved to the front (exercises the outer join and its order).
    Shared by this script and tests/test_host.py."""
    import pandas as pd
    dirs = []
    for name in ("a", "b"):
        d = os.path.join(out_root, name)
        os.makedirs(d, exist_ok=True)
        shutil.copyfile(os.path.join(src_dir, "data.json"), os.path.join(d, "data.json"))
        dirs.append(d)
    info = pd.read_csv(os.path.join(src_dir, "data.info"))
    info.to_csv(os.path.join(dirs[0], "data.info"), index=False)
    b = info.drop([0, 5, 7]).reset_index(drop=True)
    b = pd.concat([b.iloc[-10:], b.iloc[:-10]]).reset_index(drop=True)
    b.to_csv(os.path.join(dirs[1], "data.info"), index=False)
    return dirs


def replicates():
    """Reference NanopolishReplicateDS (utils/data_utils.py:341-427) on the two directories above."""
    import tempfile
    from m6anet.utils.data_utils import NanopolishReplicateDS
    src = os.path.join(REF, "m6anet", "tests", "data")
    with tempfile.TemporaryDirectory() as tmp:
        dirs = make_replicate_dirs(src, tmp)
        ds = NanopolishReplicateDS(dirs, 20, PRETRAINED_CONFIGS["HCT116_RNA002"][2], mode="Inference")
        pick = [0, 1, 3, len(ds) - 1]
        items = [ds[i] for i in pick]
        np.savez(os.path.join(HERE, "replicate_golden.npz"),
                 tx_id=np.array(ds.data_info["transcript_id"]).astype(str), tx_pos=np.array(ds.data_info["transcript_position"], dtype=np.int64),
                 n_reads=np.array(ds.data_info["n_reads"], dtype=np.int64), pick=np.array(pick),
                 **{f"feats_{j}": it[0].numpy() for j, it in enumerate(items)},
                 **{f"kmer_{j}": it[1].numpy()[0] for j, it in enumerate(items)},
                 **{f"read_id_{j}": np.array([str(r) for r in it[4]]) for j, it in enumerate(items)})
        print("replicates:", len(ds), "pooled sites")


def synthetic_inputs():
    rng = np.random.default_rng(2024)
    centre = ["".join(c) for c in __import__("itertools").product("AGT", "GA", "A", "C", "ACT")]
    seven = [a + c + b for a in "GACT" for c in centre for b in "GACT"]
    assert len(seven) == 288
    kmer_idx = np.array([[KMER_TO_INT[s[i:i + 5]] for i in range(3)] for s in seven], dtype=np.int32)
    assert len(np.unique(kmer_idx)) == 66
    n_reads = np.clip(np.round(np.exp(rng.normal(np.log(33), 0.8, size=288))), 20, 700).astype(np.int64)
    n_reads[:4] = [20, 21, 700, 32]
    read_off = np.concatenate([[0], np.cumsum(n_reads)]).astype(np.int64)
    total = int(read_off[-1])
    feats = rng.standard_normal((total, 9), dtype=np.float32)
    feats *= np.float32(1.2)
    heavy = rng.random(total) < 0.01
    feats[heavy, 0::3] = rng.uniform(8, 16, size=(int(heavy.sum()), 3)).astype(np.float32)
    np.savez(os.path.join(HERE, "synthetic_inputs.npz"), feats=feats, read_off=read_off, kmer_idx=kmer_idx,
             seven_mer=np.array(seven))
    print("synthetic:", 288, "sites", total, "reads")
    return feats, read_off, kmer_idx


def synthetic_outputs(tag, model, threshold, feats, read_off, kmer_idx, extra=None):
    n_reads = np.diff(read_off)
    kmer_rows = np.repeat(kmer_idx.astype(np.int64), n_reads, axis=0)
    p = ref_read_probs(model, feats, kmer_rows)
    site = np.zeros(len(n_reads), dtype=np.float32)
    for s in range(len(n_reads)):
        idx = sample_indices(SEED_GOLDEN, SITE_ID_BASE + s, int(n_reads[s]), N_ITERS_GOLDEN, 20)
        sl = slice(read_off[s], read_off[s + 1])
        site[s] = ref_literal_site_prob(model, feats[sl], kmer_idx[s].astype(np.int64), idx)
    mod_count = np.array([np.count_nonzero(p[read_off[s]:read_off[s + 1]] >= threshold) for s in range(len(n_reads))],
                         dtype=np.int32)
    mod_ratio = np.array([np.mean(p[read_off[s]:read_off[s + 1]] >= threshold) for s in range(len(n_reads))])
    out = dict(read_prob=p, site_prob=site, mod_count=mod_count, mod_ratio=mod_ratio, threshold=np.float64(threshold),
               seed=SEED_GOLDEN, site_id_base=SITE_ID_BASE, n_iters=N_ITERS_GOLDEN, n_samples=20)
    if extra:
        out.update(extra)
    np.savez(os.path.join(HERE, f"synthetic_{tag}.npz"), **out)
    print(tag, "p range", float(p.min()), float(p.max()), "site range", float(site.min()), float(site.max()))


def validate_golden():
    from torch.utils.data import DataLoader
    from m6anet.utils.data_utils import train_collate
    from m6anet.utils.loss_functions.loss_functions import binary_cross_entropy_loss
    from m6anet.utils.training_utils import validate
    from oracle.philox import bag_indices_mt19937

    src = os.path.join(REF, "m6anet", "tests", "data")
    shutil.copyfile(os.path.join(src, "data.info.labelled"), os.path.join(HERE, "bundled", "data.info.labelled"))
    name = "HCT116_RNA002"
    n_iters, seed = 5, 11
    flat = np.load(os.path.join(HERE, "bundled_flat.npz"))
    key = {(str(t), int(q)): i for i, (t, q) in enumerate(zip(flat["tx_id"], flat["tx_pos"]))}
    out = dict(n_iters=n_iters, seed=seed, n_samples=20)
    base_cfg = toml.load(DEFAULT_MODEL_CONFIG)
    for pool, block in (("prod", "SigmoidProdPooling"), ("mean", "SigmoidMeanPooling"), ("max", "SigmoidMaxPooling")):
        cfg = toml.loads(toml.dumps(base_cfg))
        assert cfg["block"][-1]["block_type"] == "SigmoidProdPooling"
        cfg["block"][-1]["block_type"] = block
        model = MILModel(cfg)
        model.load_state_dict(torch.load(PRETRAINED_CONFIGS[name][0], map_location="cpu"))
        for mode in ("Val", "Test"):
            ds = NanopolishDS(src, 20, PRETRAINED_CONFIGS[name][2], mode=mode)
            dl = DataLoader(ds, batch_size=16, shuffle=False, collate_fn=train_collate)   # num_workers=0: draws in site order
            np.random.seed(seed)
            res = validate(model, dl, "cpu", binary_cross_entropy_loss, n_iters)
            # replay of the legacy stream: one choice(n, 20, replace=False) per site per pass, in loader order
            rs = np.random.RandomState(seed)
            n_reads = ds.data_info["n_reads"].values
            bags = np.stack([np.stack([bag_indices_mt19937(rs, n, 20) for n in n_reads]) for _ in range(n_iters)], axis=1)
            site_index = np.array([key[(t, int(q))] for t, q in zip(ds.data_info["transcript_id"], ds.data_info["transcript_position"])])
            tag = f"{pool}_{mode}"
            out[f"{tag}_y_pred"] = np.asarray(res["y_pred"], dtype=np.float32)       # [n_iters, n_sites]
            out[f"{tag}_y_true"] = np.asarray(res["y_true"], dtype=np.int64)
            out[f"{tag}_metrics"] = np.array([res["roc_auc"], res["pr_auc"], res["avg_loss"]], dtype=np.float64)
            if pool == "prod":
                out[f"{mode}_site_index"] = site_index.astype(np.int64)              # rows of bundled_flat.npz
                out[f"{mode}_bags"] = bags.astype(np.uint16)                        # [n_sites, n_iters, 20]
            print("validate golden", tag, len(ds), "sites", out[f"{tag}_metrics"])
    np.savez_compressed(os.path.join(HERE, "validate_golden.npz"), **out)


def main():
    if sys.argv[1:] == ["validate"]:
        validate_golden()
        return
    bundled()
    validate_golden()
    replicates()
    feats, read_off, kmer_idx = synthetic_inputs()
    for name, (_, thr, _) in PRETRAINED_CONFIGS.items():
        synthetic_outputs(name, ref_model(name), thr, feats, read_off, kmer_idx)
    # signal-only topology (reference model/configs/model_configs/prod_pooling_signal.toml), seeded random weights
    torch.manual_seed(0)
    cfg = os.path.join(REF, "m6anet", "model", "configs", "model_configs", "prod_pooling_signal.toml")
    model = MILModel(toml.load(cfg))
    with torch.no_grad():   # non-trivial BatchNorm statistics
        bn = model.read_level_encoder[2].layers[1]
        bn.running_mean.normal_(0, 0.5)
        bn.running_var.uniform_(0.5, 2.0)
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_(0, 0.2)
        head = model.pooling_filter.probability_layer[0]      # spread p over (1e-4, 0.9)
        head.weight.mul_(6.0)
        head.bias.fill_(-4.0)
    model.eval()
    sd = model.state_dict()
    w = dict(w_w1=sd["read_level_encoder.2.layers.0.weight"], w_b1=sd["read_level_encoder.2.layers.0.bias"],
             w_bn_gamma=sd["read_level_encoder.2.layers.1.weight"], w_bn_beta=sd["read_level_encoder.2.layers.1.bias"],
             w_bn_mean=sd["read_level_encoder.2.layers.1.running_mean"], w_bn_var=sd["read_level_encoder.2.layers.1.running_var"],
             w_w2=sd["read_level_encoder.3.layers.0.weight"], w_b2=sd["read_level_encoder.3.layers.0.bias"],
             w_w3=sd["pooling_filter.probability_layer.0.weight"], w_b3=sd["pooling_filter.probability_layer.0.bias"])
    w = {k: v.numpy().astype(np.float32) for k, v in w.items()}
    synthetic_outputs("signal_only", model, 0.5, feats, read_off, kmer_idx, extra=w)


if __name__ == "__main__":
    main()
