from .dnn import DNN
from .deepfm import DeepFM
from .dcnv2 import DCNv2

__all__ = ["DNN", "DeepFM", "DCNv2"]
