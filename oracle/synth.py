"""TEST INFRASTRUCTURE — deterministic synthetic weights and inputs.

Weights are generated key-by-key from a seeded CPU ``torch.Generator`` so that the very
same state_dict can be rebuilt on the GPU box (no /root/reference there) and fed to the
real reference (here), the oracle restatement and the CUDA product alike.  The
distributions follow the reference initialisers:
  * conv weight/bias  U(-1/sqrt(fan_in), 1/sqrt(fan_in))   (torch Conv2d default, used by
    viewformer/models/vqgan_th.py:23-49 via torch.nn.Conv2d)
  * GroupNorm affine  gamma = 1 + 0.1 N(0,1), beta = 0.1 N(0,1)  (perturbed away from the
    torch default 1/0 so the affine path is actually exercised)
  * codebook          U(-sqrt3, sqrt3)  [D, K]               (viewformer/models/utils_th.py:17)
  * MIGT              TruncatedNormal(0.02) for wte/wpe/Conv1D (viewformer/models/migt.py:26,85,314),
                      small random biases / LN affine instead of zeros/ones so that they are tested.
"""
import math
from collections import OrderedDict

import torch


def _gen(seed):
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def _uniform(shape, bound, g):
    return (torch.rand(shape, generator=g, dtype=torch.float32) * 2 - 1) * bound


def _normal(shape, std, g):
    return torch.randn(shape, generator=g, dtype=torch.float32) * std


def _trunc_normal(shape, std, g):
    # TF TruncatedNormal: resample outside 2 sigma.  Clamp-resample approximation, deterministic.
    x = torch.randn(shape, generator=g, dtype=torch.float32)
    for _ in range(8):
        bad = x.abs() > 2
        if not bad.any():
            break
        x = torch.where(bad, torch.randn(shape, generator=g, dtype=torch.float32), x)
    return x.clamp_(-2, 2) * std


# --------------------------------------------------------------------------- VQGAN
def vqgan_param_shapes(cfg):
    """Ordered {key: shape} of the reference torch VQGAN state_dict
    (viewformer/models/vqgan_th.py:147-201, 228-289, 321-336)."""
    ch, ch_mult, nrb = cfg.ch, list(cfg.ch_mult), cfg.num_res_blocks
    nres = len(ch_mult)
    attn_res = set(cfg.attn_resolutions)
    out = OrderedDict()

    def conv(name, cout, cin, k):
        out[name + ".weight"] = (cout, cin, k, k)
        out[name + ".bias"] = (cout,)

    def norm(name, c):
        out[name + ".weight"] = (c,)
        out[name + ".bias"] = (c,)

    def resblock(name, cin, cout):
        norm(name + ".norm1", cin)
        conv(name + ".conv1", cout, cin, 3)
        norm(name + ".norm2", cout)
        conv(name + ".conv2", cout, cout, 3)
        if cin != cout:
            conv(name + ".nin_shortcut", cout, cin, 1)

    def attn(name, c):
        norm(name + ".norm", c)
        for p in ("q", "k", "v", "proj_out"):
            conv(name + "." + p, c, c, 1)

    # encoder
    conv("encoder.conv_in", ch, cfg.in_channels, 3)
    curr = cfg.image_size
    in_mult = [1] + ch_mult
    block_in = ch
    for lv in range(nres):
        block_in = ch * in_mult[lv]
        block_out = ch * ch_mult[lv]
        na = 0
        for b in range(nrb):
            resblock(f"encoder.down.{lv}.block.{b}", block_in, block_out)
            block_in = block_out
            if curr in attn_res:
                attn(f"encoder.down.{lv}.attn.{na}", block_in)
                na += 1
        if lv != nres - 1:
            conv(f"encoder.down.{lv}.downsample.conv", block_in, block_in, 3)
            curr //= 2
    resblock("encoder.mid.block_1", block_in, block_in)
    attn("encoder.mid.attn_1", block_in)
    resblock("encoder.mid.block_2", block_in, block_in)
    norm("encoder.norm_out", block_in)
    conv("encoder.conv_out", cfg.z_channels, block_in, 3)
    # decoder
    block_in = ch * ch_mult[-1]
    curr = cfg.image_size // 2 ** (nres - 1)
    conv("decoder.conv_in", block_in, cfg.z_channels, 3)
    resblock("decoder.mid.block_1", block_in, block_in)
    attn("decoder.mid.attn_1", block_in)
    resblock("decoder.mid.block_2", block_in, block_in)
    for lv in reversed(range(nres)):
        block_out = ch * ch_mult[lv]
        na = 0
        for b in range(nrb + 1):
            resblock(f"decoder.up.{lv}.block.{b}", block_in, block_out)
            block_in = block_out
            if curr in attn_res:
                attn(f"decoder.up.{lv}.attn.{na}", block_in)
                na += 1
        if lv != 0:
            conv(f"decoder.up.{lv}.upsample.conv", block_in, block_in, 3)
            curr *= 2
    norm("decoder.norm_out", block_in)
    conv("decoder.conv_out", cfg.out_ch, block_in, 3)
    # quantizer + 1x1 convs
    out["quantize.embeddings"] = (cfg.embed_dim, cfg.n_embed)
    out["quantize.ema_cluster_size_hidden"] = (cfg.n_embed,)
    out["quantize.ema_dw_hidden"] = (cfg.embed_dim, cfg.n_embed)
    out["quantize.counter"] = ()
    conv("quant_conv", cfg.embed_dim, cfg.z_channels, 1)
    conv("post_quant_conv", cfg.z_channels, cfg.embed_dim, 1)
    return out


def make_vqgan_state_dict(cfg, seed=0):
    g = _gen(seed)
    sd = OrderedDict()
    for key, shape in sorted(vqgan_param_shapes(cfg).items()):
        if key == "quantize.embeddings":
            sd[key] = _uniform(shape, math.sqrt(3.0), g)
        elif key == "quantize.counter":
            sd[key] = torch.tensor(0, dtype=torch.int64)
        elif key.startswith("quantize.ema"):
            sd[key] = torch.zeros(shape)
        elif ".norm" in key or key.endswith("norm_out.weight") or key.endswith("norm_out.bias"):
            if key.endswith(".weight"):
                sd[key] = 1.0 + _normal(shape, 0.1, g)
            else:
                sd[key] = _normal(shape, 0.1, g)
        elif key.endswith(".weight"):
            fan_in = shape[1] * shape[2] * shape[3]
            sd[key] = _uniform(shape, 1.0 / math.sqrt(fan_in), g)
        else:  # conv bias: bound uses the fan_in of the matching weight
            wshape = vqgan_param_shapes(cfg)[key[:-4] + "weight"]
            fan_in = wshape[1] * wshape[2] * wshape[3]
            sd[key] = _uniform(shape, 1.0 / math.sqrt(fan_in), g)
    return sd


# --------------------------------------------------------------------------- MIGT
def migt_param_shapes(cfg):
    """Ordered {key: shape} for the MIGT weights (viewformer/models/migt.py:76-96, 132-137,
    182-238, 288-315).  Key names follow the reference layer names (``h.<i>/attn/c_attn`` ...)."""
    d = cfg.d_model
    out = OrderedDict()
    out["wte.weight"] = (cfg.n_embeddings + 2, d)
    out["wpe.embeddings"] = (256, d)
    for n, (nx, nf) in (("pose_embedding.c_fc", (7, 2 * d)), ("pose_embedding.c_proj", (2 * d, d))):
        out[n + ".weight"] = (nx, nf)
        out[n + ".bias"] = (1, nf)
    for i in range(cfg.n_layer):
        p = f"h.{i}."
        out[p + "ln_1.gamma"] = (d,)
        out[p + "ln_1.beta"] = (d,)
        out[p + "attn.c_attn.weight"] = (d, 3 * d)
        out[p + "attn.c_attn.bias"] = (1, 3 * d)
        out[p + "attn.c_proj.weight"] = (d, d)
        out[p + "attn.c_proj.bias"] = (1, d)
        out[p + "ln_2.gamma"] = (d,)
        out[p + "ln_2.beta"] = (d,)
        out[p + "mlp.c_fc.weight"] = (d, 4 * d)
        out[p + "mlp.c_fc.bias"] = (1, 4 * d)
        out[p + "mlp.c_proj.weight"] = (4 * d, d)
        out[p + "mlp.c_proj.bias"] = (1, d)
    out["ln_f.gamma"] = (d,)
    out["ln_f.beta"] = (d,)
    for n, (nx, nf) in (("pose_classifier.c_fc", (d, 2 * d)), ("pose_classifier.c_proj", (2 * d, 7))):
        out[n + ".weight"] = (nx, nf)
        out[n + ".bias"] = (1, nf)
    return out


def make_migt_state_dict(cfg, seed=0):
    g = _gen(seed)
    sd = OrderedDict()
    for key, shape in sorted(migt_param_shapes(cfg).items()):
        if key.endswith("gamma"):
            sd[key] = 1.0 + _normal(shape, 0.05, g)
        elif key.endswith("beta") or key.endswith("bias"):
            sd[key] = _normal(shape, 0.02, g)
        else:
            sd[key] = _trunc_normal(shape, 0.02, g)
    return sd


# --------------------------------------------------------------------------- inputs
def make_images_uint8(n_scenes, n_views, size=128, seed=1234, smooth=True):
    """uint8 [B,T,H,W,3] images: low-pass filtered noise so GroupNorm/attention see
    image-like statistics; ``smooth=False`` gives plain U{0..255}."""
    g = _gen(seed)
    if not smooth:
        return torch.randint(0, 256, (n_scenes, n_views, size, size, 3), generator=g, dtype=torch.uint8)
    lo = torch.rand((n_scenes * n_views, 3, size // 8, size // 8), generator=g)
    x = torch.nn.functional.interpolate(lo, size=(size, size), mode="bilinear", align_corners=False)
    x = x + 0.08 * torch.randn(x.shape, generator=g)
    x = (x.clamp(0, 1) * 255).round().to(torch.uint8)
    return x.permute(0, 2, 3, 1).reshape(n_scenes, n_views, size, size, 3).contiguous()


def make_metric_pair(n, gt_size, gen_size, seed):
    """(ground truth uint8 [n,gt_size,gt_size,3], generated uint8 [n,gen_size,gen_size,3]) for the evaluation-metric fixtures: blocky
    images plus noise, built with integer ops only (randint / repeat / clamp) so that every machine regenerates the same bytes.
    ``gen_size`` must divide ``gt_size``; the generated image is the subsampled ground truth plus U{-20..20} noise."""
    assert gt_size % 8 == 0 and gt_size % gen_size == 0
    g = _gen(seed)
    lo = torch.randint(0, 256, (n, 8, 8, 3), generator=g)
    gt = lo.repeat_interleave(gt_size // 8, 1).repeat_interleave(gt_size // 8, 2)
    gt = (gt + torch.randint(-12, 13, gt.shape, generator=g)).clamp(0, 255)
    k = gt_size // gen_size
    gen = (gt[:, ::k, ::k] + torch.randint(-20, 21, (n, gen_size, gen_size, 3), generator=g)).clamp(0, 255)
    return gt.to(torch.uint8).contiguous(), gen.to(torch.uint8).contiguous()


def make_cameras(n_scenes, n_views, seed=4321):
    """f32 [B,T,7] = xyz ~ N(0,1) | unit quaternion with w>=0."""
    g = _gen(seed)
    xyz = torch.randn((n_scenes, n_views, 3), generator=g)
    q = torch.randn((n_scenes, n_views, 4), generator=g)
    q = q / q.norm(dim=-1, keepdim=True)
    q = q * torch.where(q[..., :1] >= 0, 1.0, -1.0)
    return torch.cat([xyz, q], -1).contiguous()


def make_codes(n_scenes, n_views, n_embed=1024, side=8, seed=99):
    g = _gen(seed)
    return torch.randint(0, n_embed, (n_scenes, n_views, side, side), generator=g, dtype=torch.int64)


def make_lookup_inputs(seed=11, D=256, K=1024):
    """Codebook E [D,K] and rows z for the nearest-neighbour fixtures (tests/golden/vq_lookup.npz):
    4096 gaussian rows, 512 adversarial near-ties (midpoints of two codes + 1e-3 noise), 64 exact codes."""
    g = _gen(seed)
    E = _uniform((D, K), math.sqrt(3.0), g)
    z = torch.randn((4096, D), generator=g)
    a = torch.randint(0, K, (512,), generator=g)
    b = torch.randint(0, K, (512,), generator=g)
    mid = 0.5 * (E[:, a] + E[:, b]).t() + 1e-3 * torch.randn((512, D), generator=g)
    return E, torch.cat([z, mid, E[:, :64].t().contiguous()], 0).contiguous()
