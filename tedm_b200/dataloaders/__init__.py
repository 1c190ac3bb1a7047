"""Input pipeline in front of the path (reference: dataloaders/JSRT.py, dataloaders/CXR14.py).

Decoding stays on the host (PIL), but pixels cross PCIe as uint8 from pinned memory on a copy stream and
are turned into the reference's fp32 tensors on the device (tedm_u8_to_unit / tedm_u8_masks_to_label)."""
from .device_loader import DeviceLoader, build_synthetic_dataloaders  # noqa: F401
from .JSRT import JSRTDataset  # noqa: F401
from .CXR14 import CXR14Dataset  # noqa: F401
