from .model import (FLAVAForPreTraining, FLAVAModel, FLAVAOutput, flava_model, flava_model_for_pretraining,  # noqa: F401
                    flava_multimodal_encoder)
from .image_encoder import flava_image_encoder, ImageTransformer  # noqa: F401
from .text_encoder import flava_text_encoder  # noqa: F401
