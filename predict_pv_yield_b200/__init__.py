"""predict_pv_yield_b200 -- B200-native (sm_100a) implementation of ONE hot path of
openclimatefix/predict_pv_yield: the Conv3d PV-yield model's train / inference step.

Layout (only what the path needs):
  csrc/            hand-written CUDA kernels + the C ABI (include/pvb200.h) -> libpvb200.so
  lib.py           ctypes binding of that ABI (fails loudly when the library is missing)
  ops.py           autograd nodes that call the ABI on torch's current stream
  models/          host-side mirror of the reference interface
                   (predict_pv_yield/models/base_model.py, models/conv3d/model.py)
  optim.py         FusedAdam (torch.optim.Optimizer surface over the Adam kernel)
  dp.py            data-parallel gradient exchange (one process per GPU, NCCL)
  batch.py, losses.py, utils.py   the few helpers of external packages the path touches
"""
__version__ = "0.1.0"
