"""Checkpoint key surgery of reference src/modeling/albef.py:208-241 for a pretrained ALBEF ``.pth``."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def interpolate_pos_embed(pos_embed_checkpoint, visual_encoder):
    """reference vit.py:193-217: bicubic resize of the patch position embeddings to the model's grid."""
    dim = pos_embed_checkpoint.shape[-1]
    num_patches = visual_encoder.patch_embed.num_patches
    extra = visual_encoder.pos_embed.shape[-2] - num_patches
    old = int((pos_embed_checkpoint.shape[-2] - extra) ** 0.5)
    new = int(num_patches ** 0.5)
    if old == new:
        return pos_embed_checkpoint
    tokens = pos_embed_checkpoint[:, extra:].reshape(-1, old, old, dim).permute(0, 3, 1, 2)
    tokens = F.interpolate(tokens, size=(new, new), mode="bicubic", align_corners=False)
    return torch.cat((pos_embed_checkpoint[:, :extra], tokens.permute(0, 2, 3, 1).flatten(1, 2)), dim=1)


def remap_albef_checkpoint(state_dict, model):
    """albef.py:208-241: 'bert.' prefixes dropped, text_encoder layers 6..11 also initialise text_decoder layers
    0..5 (the decoder starts as a copy of the multimodal half of the encoder)."""
    sd = dict(state_dict)
    sd["visual_encoder.pos_embed"] = interpolate_pos_embed(sd["visual_encoder.pos_embed"], model.visual_encoder)
    for key in list(sd.keys()):
        if "bert" in key:
            sd[key.replace("bert.", "")] = sd[key]
        if "text_encoder" in key:
            if "layer" in key:
                parts = key.split(".")
                n = int(parts[4])
                if n < 6:
                    del sd[key]
                    continue
                parts[4] = str(n - 6)
                enc_key = ".".join(parts)
            else:
                enc_key = key
            sd[enc_key.replace("text_encoder", "text_decoder")] = sd[key]
            del sd[key]
    return sd
