"""CPU restatement (plain torch) of the frozen DINO ViT feature path -- TEST INFRASTRUCTURE ONLY.

Follows third-party/zsp/zsp/method/vision_transformer_flexible.py: PatchEmbed :136-151, prepare_tokens
:214-225, interpolate_pos_encoding :192-212 (bicubic with the given scale factor (n+0.1)/n0), Block.return_qkv
:126-132, Attention.forward :85-101 (scale after q@k^T), Mlp :64-70 (exact-erf GELU), LayerNorm eps 1e-6
(:283-287); and model/module/network/dino.py:102-109 (layer-9 keys without CLS, head-major channels).
Works on a plain state dict (names as in the reference checkpoint), any embed dim / depth / heads.
"""
import math

import torch
import torch.nn.functional as F


def pos_embed_for(pos_embed, n_side_w, n_side_h):
    """pos_embed (1, 1+N0, D) -> (1, 1+w*h, D) for a w x h patch grid."""
    N0 = pos_embed.shape[1] - 1
    if n_side_w * n_side_h == N0 and n_side_w == n_side_h:
        return pos_embed
    dim = pos_embed.shape[-1]
    s0 = int(math.sqrt(N0))
    w0, h0 = n_side_w + 0.1, n_side_h + 0.1
    patch = F.interpolate(pos_embed[:, 1:].reshape(1, s0, s0, dim).permute(0, 3, 1, 2),
                          scale_factor=(w0 / s0, h0 / s0), mode='bicubic')
    assert int(w0) == patch.shape[-2] and int(h0) == patch.shape[-1]
    patch = patch.permute(0, 2, 3, 1).reshape(1, -1, dim)
    return torch.cat((pos_embed[:, :1], patch), dim=1)


def tokens(sd, x, patch):
    B, _, w, h = x.shape
    t = F.conv2d(x, sd['patch_embed.proj.weight'], sd['patch_embed.proj.bias'], stride=patch)
    t = t.flatten(2).transpose(1, 2)
    t = torch.cat((sd['cls_token'].expand(B, -1, -1), t), dim=1)
    return t + pos_embed_for(sd['pos_embed'], w // patch, h // patch)


def block(sd, i, x, heads, eps=1e-6):
    p = 'blocks.%d.' % i
    B, N, D = x.shape
    y = F.layer_norm(x, (D,), sd[p + 'norm1.weight'], sd[p + 'norm1.bias'], eps)
    qkv = F.linear(y, sd[p + 'attn.qkv.weight'], sd[p + 'attn.qkv.bias'])
    qkv = qkv.reshape(B, N, 3, heads, D // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    attn = ((q @ k.transpose(-2, -1)) * (D // heads) ** -0.5).softmax(dim=-1)
    y = (attn @ v).transpose(1, 2).reshape(B, N, D)
    x = x + F.linear(y, sd[p + 'attn.proj.weight'], sd[p + 'attn.proj.bias'])
    y = F.layer_norm(x, (D,), sd[p + 'norm2.weight'], sd[p + 'norm2.bias'], eps)
    y = F.linear(F.gelu(F.linear(y, sd[p + 'mlp.fc1.weight'], sd[p + 'mlp.fc1.bias'])),
                 sd[p + 'mlp.fc2.weight'], sd[p + 'mlp.fc2.bias'])
    return x + y, k


def keys_at_layer(sd, x, layer, heads, patch=8):
    """k of block `layer` (B, heads, 1+N, d) and the residual stream after that block."""
    t = tokens(sd, x, patch)
    k = None
    for i in range(layer + 1):
        t, k = block(sd, i, t, heads)
    return k, t


def dino_features(sd, img, layer=9, heads=6, patch=8):
    """DINO.forward of model/module/network/dino.py:102-109: (b, heads*d, h, w)."""
    k, _ = keys_at_layer(sd, img, layer, heads, patch)
    k = k[:, :, 1:, :].permute(0, 1, 3, 2)
    b, nh, d, t = k.shape
    s = int(math.sqrt(t))
    return k.reshape(b, d * nh, s, s)
