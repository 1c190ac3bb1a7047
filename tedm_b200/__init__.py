"""tedm_b200 -- B200-native (sm_100a) implementation of the data-parallel hot path of mmr12/TEDM.

Public surface mirrors the reference's modules:
    tedm_b200.models.unet_model.Unet
    tedm_b200.models.diffusion_model.DiffusionModel
    tedm_b200.models.datasetDM_model.DatasetDM
All arithmetic runs in libtedm_b200.so (hand-written CUDA, C ABI in include/tedm_b200.h).
"""
__version__ = "0.1.0"
