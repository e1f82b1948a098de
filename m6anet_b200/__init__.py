"""m6anet_b200 -- B200-native implementation of the `m6anet inference` hot path.

Host side mirrors the reference's operator surface for that path (reference m6anet/):
  m6anet_b200.inference.argparser / main / run_inference   <- scripts/inference.py, utils/inference_utils.py
  m6anet_b200.model.MILModel (TOML block list)             <- model/model.py
  m6anet_b200.constants.PRETRAINED_CONFIGS                 <- utils/constants.py
  m6anet_b200.data.NanopolishDS / NanopolishReplicateDS    <- utils/data_utils.py
The compute is one sm_100a CUDA kernel behind the C ABI in include/m6anet_b200.h
(m6anet_b200/libm6anet_b200.so); there is no CPU or PyTorch fallback for it.
"""
__version__ = "0.1.0"
