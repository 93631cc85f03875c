"""Parameter inventory of DINO ViT-S/8 (third-party/zsp/zsp/method/vision_transformer_flexible.py:153-190,
vit_small :283-287) and a deterministic synthetic state dict (no checkpoint is available offline)."""
import torch

EMBED, DEPTH, HEADS, HEAD_DIM, MLP, PATCH = 384, 12, 6, 64, 1536, 8
N_POS = 28 * 28 + 1   # pos_embed of the 224-pixel pre-training resolution


def vit_small_shapes(depth=DEPTH):
    s = {'cls_token': (1, 1, EMBED), 'pos_embed': (1, N_POS, EMBED),
         'patch_embed.proj.weight': (EMBED, 3, PATCH, PATCH), 'patch_embed.proj.bias': (EMBED,)}
    for i in range(depth):
        p = 'blocks.%d.' % i
        s.update({p + 'norm1.weight': (EMBED,), p + 'norm1.bias': (EMBED,),
                  p + 'attn.qkv.weight': (3 * EMBED, EMBED), p + 'attn.qkv.bias': (3 * EMBED,),
                  p + 'attn.proj.weight': (EMBED, EMBED), p + 'attn.proj.bias': (EMBED,),
                  p + 'norm2.weight': (EMBED,), p + 'norm2.bias': (EMBED,),
                  p + 'mlp.fc1.weight': (MLP, EMBED), p + 'mlp.fc1.bias': (MLP,),
                  p + 'mlp.fc2.weight': (EMBED, MLP), p + 'mlp.fc2.bias': (EMBED,)})
    s.update({'norm.weight': (EMBED,), 'norm.bias': (EMBED,)})
    return s


def synthetic_state_dict(seed=0, gain=2.0):
    """Seeded random weights with trained-network-like statistics: linear weights N(0, (gain*0.02)^2)
    (the reference initialises with trunc_normal(std=0.02), :179-190; gain > 1 makes attention and the
    residual stream non-trivial), LayerNorm scale 1 +- 0.1, small biases.  One generator per tensor, keyed
    by its position in the inventory, so the result does not depend on module construction order."""
    out = {}
    for idx, (name, shape) in enumerate(vit_small_shapes().items()):
        g = torch.Generator().manual_seed(seed * 1000003 + idx)
        x = torch.randn(shape, generator=g)
        if name.endswith('norm1.weight') or name.endswith('norm2.weight') or name == 'norm.weight':
            x = 1.0 + 0.1 * x
        elif name.endswith('.bias'):
            x = 0.02 * x
        elif name == 'patch_embed.proj.weight':
            x = 0.1 * x
        else:
            x = gain * 0.02 * x
        out[name] = x
    return out
