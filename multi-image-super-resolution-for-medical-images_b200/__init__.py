"""b200sr — B200-native (sm_100a) implementation of the slice-interpolation UNet hot path.

Drop-in mirror of the reference modules for that path only (reference: /root/reference/src/unet_model.py,
/root/reference/src/ModelLoader.py): same class names, constructor signatures, state_dict layout and
`load_model` entry point; the arithmetic runs in hand-written CUDA kernels behind the C ABI in include/b200sr.h.

The directory name of this package contains '-' (it is fixed by the repository layout), so it is imported
through the `b200sr` shim module at the repository root:  `import b200sr`.
"""
from ._lib import B200SRError, EXPORTED_SYMBOLS, LIB_PATH  # noqa: F401
from .unet_model import UNet, UNetBlock, UNetTrainer, MRIDataset, create_dummy_dataset  # noqa: F401
from .losses import CombinedLoss, ssim_window  # noqa: F401
from .perceptual import PerceptualLoss, VGG16Features  # noqa: F401
from .ModelLoader import load_model  # noqa: F401
from .deepcnn import DeepCNN, DeepCNNTrainer, ResidualBlock  # noqa: F401
from .progressive import (GANUNetBlock, ProgressiveUNet, ProgressiveUNetBlock, ProgressiveUNetTrainer,  # noqa: F401
                          UNetGenerator, UNetStage)
from .fastddpm import (DoubleConv, FastDDPM, FastDDPMTrainer, FastNoiseScheduler, UNet2D,  # noqa: F401
                       sinusoidal_timestep_embedding)
from .data import DevicePrefetcher, SyntheticTripletGenerator  # noqa: F401
from .metrics import compute_metrics  # noqa: F401

__version__ = "0.1.0"
