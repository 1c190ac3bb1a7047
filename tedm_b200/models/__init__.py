from .unet_model import Unet  # noqa: F401
from .diffusion_model import DiffusionModel  # noqa: F401
from .datasetDM_model import DatasetDM, tedm_classifier  # noqa: F401
