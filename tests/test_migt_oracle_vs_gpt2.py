"""Hardening of the (unpinned) MIGT oracle: the reference transformer is a GPT-2 derivative (models/migt.py:59-96 Conv1D / MLP,
:182-238 attention + pre-LN block are the HF TFGPT2 layers with the (v,q,k) split, no 1/sqrt(d) scale and block-causal masking).
With ONE token per view (token_image_size = 1) block-causal attention degenerates to ordinary causal attention, so the oracle's
`block()` must reproduce `transformers`' torch GPT2Block (scale_attn_weights=False, exact-erf GELU, LayerNorm eps 1e-5) once the
c_attn columns are permuted from (v,q,k) to HF's (q,k,v).  An independent implementation agreeing to 1e-5 pins Conv1D / LN / GELU /
residual wiring / softmax attention of the restatement to third-party code; the multi-token block mask and the extra streams stay
covered by the structural invariants of tests/test_oracle_pinned.py."""
import pytest
import torch

from oracle import migt_oracle as mo

transformers = pytest.importorskip("transformers")


def _gpt2_block(d, n_head, sd, prefix):
    from transformers import GPT2Config
    from transformers.models.gpt2.modeling_gpt2 import GPT2Block
    cfg = GPT2Config(n_embd=d, n_head=n_head, n_layer=1, n_positions=64, activation_function="gelu", layer_norm_epsilon=1e-5,
                     scale_attn_weights=False, attn_pdrop=0.0, resid_pdrop=0.0, embd_pdrop=0.0)
    cfg._attn_implementation = "eager"
    blk = GPT2Block(cfg, layer_idx=0).eval()
    w = sd[prefix + "attn.c_attn.weight"]                      # [d, 3d], columns v | q | k  (migt.py:207-213)
    b = sd[prefix + "attn.c_attn.bias"].reshape(-1)
    with torch.no_grad():
        blk.attn.c_attn.weight.copy_(torch.cat([w[:, d:2 * d], w[:, 2 * d:], w[:, :d]], 1))        # -> q | k | v
        blk.attn.c_attn.bias.copy_(torch.cat([b[d:2 * d], b[2 * d:], b[:d]]))
        blk.attn.c_proj.weight.copy_(sd[prefix + "attn.c_proj.weight"]); blk.attn.c_proj.bias.copy_(sd[prefix + "attn.c_proj.bias"].reshape(-1))
        blk.mlp.c_fc.weight.copy_(sd[prefix + "mlp.c_fc.weight"]); blk.mlp.c_fc.bias.copy_(sd[prefix + "mlp.c_fc.bias"].reshape(-1))
        blk.mlp.c_proj.weight.copy_(sd[prefix + "mlp.c_proj.weight"]); blk.mlp.c_proj.bias.copy_(sd[prefix + "mlp.c_proj.bias"].reshape(-1))
        blk.ln_1.weight.copy_(sd[prefix + "ln_1.gamma"]); blk.ln_1.bias.copy_(sd[prefix + "ln_1.beta"])
        blk.ln_2.weight.copy_(sd[prefix + "ln_2.gamma"]); blk.ln_2.bias.copy_(sd[prefix + "ln_2.beta"])
    return blk


@pytest.mark.parametrize("d,n_head,T", [(64, 4, 12), (96, 3, 7)])
def test_oracle_block_equals_hf_gpt2_block(d, n_head, T):
    g = torch.Generator().manual_seed(d + T)
    p = "h.0."
    sd = {}
    for name, (nx, nf) in (("attn.c_attn", (d, 3 * d)), ("attn.c_proj", (d, d)), ("mlp.c_fc", (d, 4 * d)), ("mlp.c_proj", (4 * d, d))):
        sd[p + name + ".weight"] = torch.randn(nx, nf, generator=g) * 0.08
        sd[p + name + ".bias"] = torch.randn(1, nf, generator=g) * 0.05
    for ln in ("ln_1", "ln_2"):
        sd[p + ln + ".gamma"] = 1 + 0.1 * torch.randn(d, generator=g)
        sd[p + ln + ".beta"] = 0.1 * torch.randn(d, generator=g)
    x = torch.randn(2, T, 1, d, generator=g)                   # [B, T views, L = 1 token, d]
    with torch.no_grad():
        got = mo.block(sd, p, [x], n_head)[0].reshape(2, T, d)
        blk = _gpt2_block(d, n_head, sd, p)
        # transformers >= 5 builds the causal mask at model level: hand the block an explicit additive mask
        causal = torch.full((T, T), torch.finfo(torch.float32).min).triu(1)[None, None].expand(2, 1, T, T)
        out = blk(x.reshape(2, T, d), attention_mask=causal)
        want = out[0] if isinstance(out, (tuple, list)) else out
    err = float((got - want).abs().max())
    print(f"[oracle block vs HF GPT2Block d={d} heads={n_head} T={T}] max abs diff {err:.2e}")
    assert err < 2e-5


@pytest.mark.parametrize("d,n_head,T,L", [(64, 4, 5, 4), (96, 3, 3, 8)])
def test_oracle_block_causal_mask_equals_hf_gpt2_block_with_block_mask(d, n_head, T, L):
    """L tokens per view: the oracle's block-causal attention (branching_attention.py:41-61: a token sees every token of its own and of
    earlier views, masked logits -1e4) against HF's GPT2Block handed the same visibility as an additive mask — pins the mask SHAPE of the
    restatement (view >= view, not token >= token) to an independent implementation."""
    g = torch.Generator().manual_seed(d + T + L)
    p = "h.0."
    sd = {}
    for name, (nx, nf) in (("attn.c_attn", (d, 3 * d)), ("attn.c_proj", (d, d)), ("mlp.c_fc", (d, 4 * d)), ("mlp.c_proj", (4 * d, d))):
        sd[p + name + ".weight"] = torch.randn(nx, nf, generator=g) * 0.08
        sd[p + name + ".bias"] = torch.randn(1, nf, generator=g) * 0.05
    for ln in ("ln_1", "ln_2"):
        sd[p + ln + ".gamma"] = 1 + 0.1 * torch.randn(d, generator=g)
        sd[p + ln + ".beta"] = 0.1 * torch.randn(d, generator=g)
    x = torch.randn(2, T, L, d, generator=g)
    S = T * L
    view = torch.arange(S) // L
    visible = view[:, None] >= view[None, :]
    add_mask = torch.where(visible, 0.0, torch.finfo(torch.float32).min)[None, None].expand(2, 1, S, S)
    with torch.no_grad():
        got = mo.block(sd, p, [x], n_head)[0].reshape(2, S, d)
        out = _gpt2_block(d, n_head, sd, p)(x.reshape(2, S, d), attention_mask=add_mask)
        want = out[0] if isinstance(out, (tuple, list)) else out
    err = float((got - want).abs().max())
    print(f"[oracle block-causal vs HF GPT2Block + block mask d={d} heads={n_head} T={T} L={L}] max abs diff {err:.2e}")
    assert err < 2e-5
    # and it is NOT ordinary token-causal attention: with the token-causal mask HF's output differs
    with torch.no_grad():
        tok = torch.full((S, S), torch.finfo(torch.float32).min).triu(1)[None, None].expand(2, 1, S, S)
        out_t = _gpt2_block(d, n_head, sd, p)(x.reshape(2, S, d), attention_mask=tok)
        out_t = out_t[0] if isinstance(out_t, (tuple, list)) else out_t
    assert float((got - out_t).abs().max()) > 1e-3
