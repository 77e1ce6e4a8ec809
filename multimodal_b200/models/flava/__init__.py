from .model import FLAVAModel, FLAVAOutput, flava_model, flava_multimodal_encoder  # noqa: F401
from .image_encoder import flava_image_encoder, ImageTransformer  # noqa: F401
from .text_encoder import flava_text_encoder  # noqa: F401
