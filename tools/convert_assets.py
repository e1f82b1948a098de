#!/usr/bin/env python
"""Convert the reference's pretrained assets into the neutral .npz files this package ships.

Run once in the build container (needs /root/reference, torch, joblib):
    python tools/convert_assets.py [--reference /root/reference]

Inputs  (reference m6anet/utils/constants.py:10-27):
    m6anet/model/model_states/*.pt        torch state_dicts saved from CUDA
    m6anet/model/norm_factors/*.joblib    dict 5-mer -> (mean[3], std[3]) float64
Outputs (m6anet_b200/assets/):
    model_states/<stem>.npz   raw, UNFOLDED float32 tensors: emb w1 b1 bn_gamma bn_beta bn_mean
                              bn_var bn_eps w2 b2 w3 b3  (BatchNorm folding happens at load time)
    norm_factors/<stem>.npz   kmers <U5 [K] (sorted), mean f64 [K,3], std f64 [K,3]
The values are bit-identical to the reference files; only the container format changes, so the
GPU box (which has no /root/reference) can load them without torch.load/joblib.
"""
import argparse
import os

import numpy as np

KEYMAP = {
    "emb": "read_level_encoder.1.embedding_layer.weight",
    "w1": "read_level_encoder.3.layers.0.weight",
    "b1": "read_level_encoder.3.layers.0.bias",
    "bn_gamma": "read_level_encoder.3.layers.1.weight",
    "bn_beta": "read_level_encoder.3.layers.1.bias",
    "bn_mean": "read_level_encoder.3.layers.1.running_mean",
    "bn_var": "read_level_encoder.3.layers.1.running_var",
    "w2": "read_level_encoder.4.layers.0.weight",
    "b2": "read_level_encoder.4.layers.0.bias",
    "w3": "pooling_filter.probability_layer.0.weight",
    "b3": "pooling_filter.probability_layer.0.bias",
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "m6anet_b200", "assets"))
    args = ap.parse_args()
    import joblib
    import torch

    src = os.path.join(args.reference, "m6anet", "model")
    os.makedirs(os.path.join(args.out, "model_states"), exist_ok=True)
    os.makedirs(os.path.join(args.out, "norm_factors"), exist_ok=True)
    for fn in sorted(os.listdir(os.path.join(src, "model_states"))):
        if not fn.endswith(".pt"):
            continue
        sd = torch.load(os.path.join(src, "model_states", fn), map_location="cpu")
        arrs = {k: sd[v].detach().cpu().numpy().astype(np.float32) for k, v in KEYMAP.items()}
        arrs["bn_eps"] = np.float64(1e-5)
        out = os.path.join(args.out, "model_states", fn[:-3] + ".npz")
        np.savez(out, **arrs)
        print("wrote", out, {k: v.shape for k, v in arrs.items()})
    for fn in sorted(os.listdir(os.path.join(src, "norm_factors"))):
        if not fn.endswith(".joblib"):
            continue
        nd = joblib.load(os.path.join(src, "norm_factors", fn))
        kmers = np.array(sorted(nd.keys()))
        mean = np.stack([np.asarray(nd[k][0], dtype=np.float64) for k in kmers])
        std = np.stack([np.asarray(nd[k][1], dtype=np.float64) for k in kmers])
        out = os.path.join(args.out, "norm_factors", fn[: -len(".joblib")] + ".npz")
        np.savez(out, kmers=kmers, mean=mean, std=std)
        print("wrote", out, kmers.shape, mean.shape, std.shape)


if __name__ == "__main__":
    main()
