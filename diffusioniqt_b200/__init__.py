"""diffusioniqt_b200: B200-native sampling hot path of DiffusionIQT.

Public names mirror /root/reference/imagen_pytorch3D.py so that
    from diffusioniqt_b200 import Unet, SRUnet256, NullUnet, Imagen, ElucidatedImagen
is a drop-in for the reference's sampling path.  The arithmetic lives in libdiqt_b200.so
(include/diqt.h); importing this package does not need a GPU, running it does.
"""
from .unet import Unet, Unet3D, SRUnet256, SRUnet1024, BaseUnet64, NullUnet  # noqa: F401
from .imagen import Imagen, GaussianDiffusionContinuousTimes  # noqa: F401
from .elucidated import ElucidatedImagen, Hparams  # noqa: F401
from .trainer import ImagenTrainer  # noqa: F401

__version__ = "0.1.0"
