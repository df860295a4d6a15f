"""Model configs (reference src/configs/model_configs.py:9-15, 40-90): the ViLT entry and the two ALBEF entries."""
from ..modeling.albef import ALBEFWrapper, convert_batch_to_albef_input_dict
from ..modeling.albef_model import CONFIG_BERT
from ..modeling.vilt import ViltEncoderWrapper, convert_batch_to_vilt_input_dict

ALLOWED_CL_ENCODERS = ["vilt", "albef_distill", "albef_no_distill"]        # reference modeling/__init__.py

vilt_config = {
    "encoder_dim": 768,
    "visual_input_type": "pil-image",
    "encoder_class": ViltEncoderWrapper,
    "batch2inputs_converter": convert_batch_to_vilt_input_dict,
    "encoder_name": "ViLT",
}

config_bert = dict(CONFIG_BERT)                                              # model_configs.py:40-60

albef_no_distill_config = {                                                  # model_configs.py:62-72
    "text_encoder": "bert-base-uncased",
    "text_decoder": "bert-base-uncased",
    "image_res": 384,
    "visual_input_type": "pil-image",
    "bert_config": config_bert,
    "batch2inputs_converter": convert_batch_to_albef_input_dict,
    "distill": False,
    "encoder_class": ALBEFWrapper,
    "encoder_name": "albef_no_distill",
}

albef_distill_config = dict(albef_no_distill_config, distill=True, encoder_name="albef_distill")   # :74-84

model_configs = {"vilt": vilt_config, "albef_distill": albef_distill_config,
                 "albef_no_distill": albef_no_distill_config}
