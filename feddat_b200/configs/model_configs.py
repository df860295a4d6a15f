"""Model configs (reference src/configs/model_configs.py:9-15,86-90), ViLT entry."""
from ..modeling.vilt import ViltEncoderWrapper, convert_batch_to_vilt_input_dict

ALLOWED_CL_ENCODERS = ["vilt"]

vilt_config = {
    "encoder_dim": 768,
    "visual_input_type": "pil-image",
    "encoder_class": ViltEncoderWrapper,
    "batch2inputs_converter": convert_batch_to_vilt_input_dict,
    "encoder_name": "ViLT",
}

model_configs = {"vilt": vilt_config}
