"""TEST INFRASTRUCTURE — torch-CPU functional restatement of the reference torch VQGAN.

Follows (file:line in /root/reference):
  viewformer/models/vqgan_th.py:11-17   swish, GroupNorm(32, eps=1e-6)
  viewformer/models/vqgan_th.py:20-49   Upsample (nearest x2 + conv3x3), Downsample (pad (0,1,0,1) + stride-2 conv)
  viewformer/models/vqgan_th.py:52-90   ResnetBlock
  viewformer/models/vqgan_th.py:93-144  AttnBlock (single head, scale C^-0.5)
  viewformer/models/vqgan_th.py:203-225 Encoder.forward ; :291-318 Decoder.forward
  viewformer/models/vqgan_th.py:379-398 encode / decode / decode_code / forward
  viewformer/models/utils_th.py:32-72   QuantizeEMA.forward (eval + train branch), embed_code
  viewformer/models/utils_th.py:93-120  Quantize.forward (beta = 0.25 commit loss)

Pinned: tests/test_oracle_pinned.py checks this restatement against the *real* reference module
(imported by oracle/ref_loader.py, container only) and against tests/golden/vqgan_*.npz, which were
produced by the real reference (oracle/make_golden.py).

It is written against a plain ``state_dict`` with the reference's key names, not as a module mirror.
"""
import torch
import torch.nn.functional as F


def swish(x):
    return x * torch.sigmoid(x)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _conv(sd, p, x, stride=1, padding=0):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resblock(sd, p, x):
    h = _conv(sd, p + ".conv1", swish(_gn(sd, p + ".norm1", x)), padding=1)
    h = _conv(sd, p + ".conv2", swish(_gn(sd, p + ".norm2", h)), padding=1)
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x)
    return x + h


def attnblock(sd, p, x):
    h = _gn(sd, p + ".norm", x)
    q, k, v = (_conv(sd, p + "." + n, h) for n in ("q", "k", "v"))
    b, c, hh, ww = q.shape
    q = q.reshape(b, c, hh * ww).permute(0, 2, 1)
    k = k.reshape(b, c, hh * ww)
    w = torch.bmm(q, k) * (int(c) ** (-0.5))
    w = F.softmax(w, dim=2)
    v = v.reshape(b, c, hh * ww)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + ".proj_out", h)


def _levels(cfg):
    nres = len(cfg.ch_mult)
    res = [cfg.image_size // 2 ** i for i in range(nres)]
    return nres, res


def encoder(sd, cfg, x):
    nres, res = _levels(cfg)
    h = _conv(sd, "encoder.conv_in", x, padding=1)
    for lv in range(nres):
        na = 0
        for b in range(cfg.num_res_blocks):
            h = resblock(sd, f"encoder.down.{lv}.block.{b}", h)
            if res[lv] in cfg.attn_resolutions:
                h = attnblock(sd, f"encoder.down.{lv}.attn.{na}", h)
                na += 1
        if lv != nres - 1:
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = _conv(sd, f"encoder.down.{lv}.downsample.conv", h, stride=2)
    h = resblock(sd, "encoder.mid.block_1", h)
    h = attnblock(sd, "encoder.mid.attn_1", h)
    h = resblock(sd, "encoder.mid.block_2", h)
    h = swish(_gn(sd, "encoder.norm_out", h))
    return _conv(sd, "encoder.conv_out", h, padding=1)


def decoder(sd, cfg, z):
    nres, res = _levels(cfg)
    h = _conv(sd, "decoder.conv_in", z, padding=1)
    h = resblock(sd, "decoder.mid.block_1", h)
    h = attnblock(sd, "decoder.mid.attn_1", h)
    h = resblock(sd, "decoder.mid.block_2", h)
    for lv in reversed(range(nres)):
        na = 0
        for b in range(cfg.num_res_blocks + 1):
            h = resblock(sd, f"decoder.up.{lv}.block.{b}", h)
            if res[lv] in cfg.attn_resolutions:
                h = attnblock(sd, f"decoder.up.{lv}.attn.{na}", h)
                na += 1
        if lv != 0:
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"decoder.up.{lv}.upsample.conv", h, padding=1)
    h = swish(_gn(sd, "decoder.norm_out", h))
    return _conv(sd, "decoder.conv_out", h, padding=1)


def embed_code(embeddings, ids):
    """utils_th.py:70-72 — ids [N,h,w] -> [N,D,h,w]."""
    return F.embedding(ids, embeddings.t()).permute(0, 3, 1, 2).contiguous()


def vq_lookup(embeddings, flat):
    """utils_th.py:34-41 — expanded-form fp32 distance; (-dist).max(1) => first index wins ties."""
    dist = flat.pow(2).sum(1, keepdim=True) - 2 * flat @ embeddings + embeddings.pow(2).sum(0, keepdim=True)
    return (-dist).max(1)[1]


def quantize_ema(sd, z, training=False, world_sums=None, decay=0.99, eps=1e-5):
    """utils_th.py:32-68.  Returns (quantize, diff, ids[, new_buffers when training]).

    ``world_sums`` optionally replaces the all_reduce at :50-52: a callable mapping
    (embed_onehot_sum, embed_sum) -> their sums over ranks."""
    E = sd["quantize.embeddings"]
    x = z.permute(0, 2, 3, 1)
    flat = x.reshape(-1, x.size(-1))
    ind = vq_lookup(E, flat)
    onehot = F.one_hot(ind, E.shape[1]).type(flat.dtype)
    ids = ind.view(*x.shape[:-1])
    q = embed_code(E, ids)
    new = None
    if training:
        onehot_sum = onehot.sum(0)
        embed_sum = flat.transpose(0, 1) @ onehot
        if world_sums is not None:
            onehot_sum, embed_sum = world_sums(onehot_sum, embed_sum)
        cs = sd["quantize.ema_cluster_size_hidden"].clone()
        dw = sd["quantize.ema_dw_hidden"].clone()
        counter = sd["quantize.counter"].clone()
        cs.add_(onehot_sum - cs, alpha=1 - decay)
        dw.add_(embed_sum - dw, alpha=1 - decay)
        counter.add_(1)
        corr = 1.0 - torch.pow(torch.tensor(decay), counter)
        ema_cs = cs / corr
        ema_dw = dw / corr
        n = ema_cs.sum()
        cluster = (ema_cs + eps) / (n + E.shape[1] * eps) * n
        new = {"quantize.ema_cluster_size_hidden": cs, "quantize.ema_dw_hidden": dw,
               "quantize.counter": counter, "quantize.embeddings": ema_dw / cluster.unsqueeze(0)}
    diff = (q.detach() - z).pow(2).mean()
    q = z + (q - z).detach()
    return (q, diff, ids, new) if training else (q, diff, ids)


def quantize_commit(embeddings, z, beta=0.25):
    """utils_th.py:93-120 (`Quantize`, the beta=0.25 two-term VQ+commit loss)."""
    x = z.permute(0, 2, 3, 1)
    flat = x.reshape(-1, x.size(-1))
    ids = vq_lookup(embeddings, flat).view(*x.shape[:-1])
    q = embed_code(embeddings, ids)
    loss = torch.mean((q.detach() - z).pow(2)) + beta * torch.mean((q - z.detach()).pow(2))
    return z + (q - z).detach(), loss, ids


def encode(sd, cfg, x, return_pre_quant=False):
    """vqgan_th.py:379-383 (eval mode) — x NCHW f32 in [-1,1] -> (quant, diff, codes int64 [N,h,w])."""
    h = _conv(sd, "quant_conv", encoder(sd, cfg, x))
    q, diff, ids = quantize_ema(sd, h)
    return (q, diff, ids, h) if return_pre_quant else (q, diff, ids)


def decode(sd, cfg, quant):
    return decoder(sd, cfg, _conv(sd, "post_quant_conv", quant))


def decode_code(sd, cfg, codes):
    return decode(sd, cfg, embed_code(sd["quantize.embeddings"], codes))


def forward(sd, cfg, x):
    q, diff, ids = encode(sd, cfg, x)
    return decode(sd, cfg, q), diff, q, ids


def compute_loss(cfg, codebook_loss, inputs, recon):
    """vqgan_th.py:400-408 with perceptual_weight == 0 (LPIPS weights unavailable offline)."""
    rec = torch.abs(inputs - recon).mean()
    return rec + cfg.codebook_weight * codebook_loss.mean()
