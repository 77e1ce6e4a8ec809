from .coca_model import (  # noqa: F401
    coca_for_pretraining, coca_vit, coca_vit_b_32, coca_vit_l_14, CoCaForPretraining, CoCaModel, MultimodalOutput,
)
