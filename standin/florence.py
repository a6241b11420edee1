"""Seeded Florence-2-base stand-in through the installed ``transformers`` classes: the CPU oracle's caption network AND
the definition of the seeded stand-in checkpoint (test/benchmark infrastructure, not product; see standin/__init__.py).

The reference loads ``microsoft/Florence-2-base`` remote code + fine-tuned ``icon_caption`` safetensors
(ref:util/utils.py:62-68); neither is available offline.  ``transformers`` 5.5 ships the same network natively
(``models/florence2``: DaViT tower + projector + BART 6+6, image tokens first then prompt tokens), so the oracle is
that library class with seeded weights: "parity unpinned" at the weight level, pinned at the architecture level.
Generation settings restate the Florence-2-base checkpoint's ``generation_config.json`` as recalled in
SURVEY.md §8a F5 (forced BOS 0, forced EOS 2, no_repeat_ngram_size 3, eos 2, pad 1, decoder_start 2); to be
re-verified against a real ``weights/icon_caption_florence`` directory when one is available.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

VOCAB = 51290          # 51289 BART ids + the native image placeholder id 51289
IMAGE_TOKEN = 51289
PROMPT_IDS = [0, 2264, 473, 5, 2274, 6190, 116, 2]   # "<s>What does the image describe?</s>" (ref:util/utils.py:109-110)
GEN = dict(forced_bos_token_id=0, forced_eos_token_id=2, no_repeat_ngram_size=3, eos_token_id=2, pad_token_id=1,
           bos_token_id=0, decoder_start_token_id=2)
IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def florence_config():
    from transformers import Florence2Config

    text = dict(model_type="bart", vocab_size=VOCAB, d_model=768, encoder_layers=6, decoder_layers=6,
                encoder_attention_heads=12, decoder_attention_heads=12, encoder_ffn_dim=3072, decoder_ffn_dim=3072,
                max_position_embeddings=1024, activation_function="gelu", scale_embedding=False, dropout=0.0,
                attention_dropout=0.0, activation_dropout=0.0, bos_token_id=0, eos_token_id=2, pad_token_id=1,
                decoder_start_token_id=2, forced_eos_token_id=2, tie_word_embeddings=True)
    cfg = Florence2Config(text_config=text, vision_config=dict(projection_dim=768, drop_path_rate=0.0),
                          image_token_id=IMAGE_TOKEN)
    return cfg


def florence_standin(seed: int = 0):
    """Florence2ForConditionalGeneration (eval, fp32, CPU) with seeded unit-gain weights."""
    from transformers import Florence2ForConditionalGeneration

    cfg = florence_config()
    cfg._attn_implementation = "eager"
    model = Florence2ForConditionalGeneration(cfg).eval()
    g = torch.Generator().manual_seed(1234 + seed)
    with torch.no_grad():
        for name, mod in model.named_modules():
            if isinstance(mod, nn.Linear):
                if name == "lm_head":
                    continue   # tied to the shared embedding
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) / math.sqrt(mod.in_features))
                if mod.bias is not None:
                    mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
            elif isinstance(mod, nn.Conv2d):
                fan_in = mod.in_channels // mod.groups * mod.kernel_size[0] * mod.kernel_size[1]
                mod.weight.copy_(torch.randn(mod.weight.shape, generator=g) / math.sqrt(fan_in))
                if mod.bias is not None:
                    mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
            elif isinstance(mod, nn.LayerNorm):
                mod.weight.copy_(1.0 + 0.1 * torch.randn(mod.weight.shape, generator=g))
                mod.bias.copy_(0.1 * torch.randn(mod.bias.shape, generator=g))
            elif isinstance(mod, nn.Embedding):
                std = 0.05 if mod.num_embeddings == VOCAB else 0.5
                mod.weight.copy_(std * torch.randn(mod.weight.shape, generator=g))
        model.tie_weights()
    model.generation_config.update(**GEN, num_beams=1, do_sample=False)
    model.config.name_or_path = "seeded/florence2-standin"   # ref:util/utils.py:109 looks for 'florence'
    return model


def pixel_values_from_u8(crops_u8: torch.Tensor) -> torch.Tensor:
    """[K,64,64,3] u8 -> [K,3,64,64] f32: rescale 1/255 then ImageNet normalise, the CLIP image processor's
    ``do_resize=False`` branch used at ref:util/utils.py:121 (Florence-2 preprocessor_config)."""
    x = crops_u8.to(torch.float32).permute(0, 3, 1, 2) * (1.0 / 255.0)
    mean = torch.tensor(IMAGENET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(IMAGENET_STD).view(1, 3, 1, 1)
    return (x - mean) / std


def input_ids_for(n_crops: int, n_image_tokens: int = 5) -> torch.Tensor:
    """image placeholders first, then the prompt (hf:models/florence2/processing_florence2.py:182-187)."""
    row = [IMAGE_TOKEN] * n_image_tokens + PROMPT_IDS
    return torch.tensor([row] * n_crops, dtype=torch.long)




def to_remote_code_names(sd):
    """Native ``transformers`` Florence-2 parameter names -> the microsoft/Florence-2 remote-code names (SURVEY.md §8c,
    recalled layout) a real ``icon_caption_florence/model.safetensors`` uses.  Fixture for the loader tests: the inverse
    of ``omniparser_b200.caption.rename_remote_code``, written independently of it."""
    out = {}
    for k, v in sd.items():
        if k.startswith("model.vision_tower.convs."):
            nk = k[len("model."):].replace(".conv.weight", ".proj.weight").replace(".conv.bias", ".proj.bias")
        elif k.startswith("model.vision_tower.blocks."):
            nk = k[len("model."):]
            kind = "window_attn" if ".spatial_block." in nk else "channel_attn"
            nk = nk.replace(".norm1.", f".{kind}.norm.").replace(".norm2.", ".ffn.norm.")
            nk = nk.replace(f".{kind}.qkv.", f".{kind}.fn.qkv.").replace(f".{kind}.proj.", f".{kind}.fn.proj.")
            nk = nk.replace(".conv1.", ".conv1.fn.dw.").replace(".conv2.", ".conv2.fn.dw.")
            nk = nk.replace(".ffn.fc1.", ".ffn.fn.net.fc1.").replace(".ffn.fc2.", ".ffn.fn.net.fc2.")
        elif k == "model.multi_modal_projector.image_projection.weight":
            nk, v = "image_projection", v.t().contiguous()
        elif k.startswith("model.multi_modal_projector.image_proj_norm."):
            nk = k[len("model.multi_modal_projector."):]
        elif k.startswith("model.multi_modal_projector.image_position_embed."):
            nk = "image_pos_embed." + k[len("model.multi_modal_projector.image_position_embed."):]
        elif k.startswith("model.multi_modal_projector.visual_temporal_embed."):
            nk = k[len("model.multi_modal_projector."):]
        elif k.startswith("model.language_model."):
            nk = "language_model.model." + k[len("model.language_model."):]
        elif k == "lm_head.weight":
            nk = "language_model.lm_head.weight"
        else:
            raise KeyError(k)
        out[nk] = v
    out["language_model.final_logits_bias"] = torch.zeros((1, sd["lm_head.weight"].shape[0]))
    return out


def export_remote_code_dir(model, path) -> None:
    """Write ``model.safetensors`` (remote-code names) + ``config.json`` + ``generation_config.json`` the way
    ``weights/icon_caption_florence`` is laid out (ref:README.md:45-46)."""
    import json
    import os
    from safetensors.torch import save_file
    os.makedirs(str(path), exist_ok=True)
    sd = {k: v.detach().clone().contiguous() for k, v in to_remote_code_names(model.state_dict()).items()}
    save_file(sd, os.path.join(str(path), "model.safetensors"))
    with open(os.path.join(str(path), "generation_config.json"), "w") as f:
        json.dump(dict(GEN, num_beams=3, early_stopping=True), f)
    with open(os.path.join(str(path), "config.json"), "w") as f:
        json.dump({"model_type": "florence2", "text_config": {"decoder_start_token_id": 2}}, f)
