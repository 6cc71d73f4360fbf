"""lsdm_b200 -- B200-native (sm_100a) implementation of LSDM's multi-conditional denoising path.

Python surface = the reference's (``util.model_util``, ``model.sdm``, ``diffusion.*``); arithmetic = hand-written
CUDA in ``liblsdm_b200.so`` behind the C ABI of ``include/lsdm_b200.h``.  See DESIGN.md / INTEGRATION.md.
"""
__version__ = "0.1"
