"""Third-party cross-check of the MIGT oracle (its pin to the reference's own code is tests/test_reference_on_shim.py): the reference transformer is a GPT-2 derivative (models/migt.py:59-96 Conv1D / MLP,
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


def test_oracle_forward_body_equals_hf_gpt2_model():
    """Whole single-stream body (migt.py:338-417 with use_localization off): N blocks, ln_f and the tied LM head sliced to n_embeddings,
    against transformers' GPT2Model fed the oracle's own input embedding (token + per-view position + pose embedding) through
    `inputs_embeds` with its wpe zeroed and the block-causal visibility as the attention mask.  Pins the layer stacking, the final
    LayerNorm and the `[..., :n_embeddings]` logits of the restatement to the third-party model."""
    from transformers import GPT2Config, GPT2Model
    from oracle import synth
    from viewformer_b200.config import MIGTConfig
    cfg = MIGTConfig(n_layer=3, n_head=4, d_model=64, sequence_size=5, n_embeddings=40, token_image_size=2, localization_weight="0")
    sd = synth.make_migt_state_dict(cfg, 3)
    B, T, L, d, V = 2, 5, 4, 64, 40
    ids = torch.randint(0, V + 1, (B, T, 2, 2), generator=torch.Generator().manual_seed(1))      # includes the mask token V
    poses = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(B, T, seed=2))[0])
    with torch.no_grad():
        out = mo.forward(sd, cfg, dict(input_ids=ids, poses=poses), use_localization=False)
        # the oracle's input embedding, restated from its three terms (migt.py:362-386)
        emb = sd["wte.weight"][ids.reshape(B, T, L)] + sd["wpe.embeddings"][:L][None, None] + \
            mo.mlp(sd, "pose_embedding", mo.pose_model_input(cfg, poses.float())).unsqueeze(-2)
        hf_cfg = GPT2Config(n_embd=d, n_head=4, n_layer=3, n_positions=T * L, vocab_size=V + 2, activation_function="gelu",
                            layer_norm_epsilon=1e-5, scale_attn_weights=False, attn_pdrop=0.0, resid_pdrop=0.0, embd_pdrop=0.0)
        hf_cfg._attn_implementation = "eager"
        hf = GPT2Model(hf_cfg).eval()
        hf.wpe.weight.zero_()
        for i, blk in enumerate(hf.h):
            p = f"h.{i}."
            w, b = sd[p + "attn.c_attn.weight"], sd[p + "attn.c_attn.bias"].reshape(-1)
            blk.attn.c_attn.weight.copy_(torch.cat([w[:, d:2 * d], w[:, 2 * d:], w[:, :d]], 1)); blk.attn.c_attn.bias.copy_(torch.cat([b[d:2 * d], b[2 * d:], b[:d]]))
            blk.attn.c_proj.weight.copy_(sd[p + "attn.c_proj.weight"]); blk.attn.c_proj.bias.copy_(sd[p + "attn.c_proj.bias"].reshape(-1))
            blk.mlp.c_fc.weight.copy_(sd[p + "mlp.c_fc.weight"]); blk.mlp.c_fc.bias.copy_(sd[p + "mlp.c_fc.bias"].reshape(-1))
            blk.mlp.c_proj.weight.copy_(sd[p + "mlp.c_proj.weight"]); blk.mlp.c_proj.bias.copy_(sd[p + "mlp.c_proj.bias"].reshape(-1))
            blk.ln_1.weight.copy_(sd[p + "ln_1.gamma"]); blk.ln_1.bias.copy_(sd[p + "ln_1.beta"])
            blk.ln_2.weight.copy_(sd[p + "ln_2.gamma"]); blk.ln_2.bias.copy_(sd[p + "ln_2.beta"])
        hf.ln_f.weight.copy_(sd["ln_f.gamma"]); hf.ln_f.bias.copy_(sd["ln_f.beta"])
        S = T * L
        view = torch.arange(S) // L
        add_mask = torch.where(view[:, None] >= view[None, :], 0.0, torch.finfo(torch.float32).min)[None, None].expand(B, 1, S, S)
        hid = hf(inputs_embeds=emb.reshape(B, S, d), attention_mask=add_mask).last_hidden_state
        want_logits = (hid @ sd["wte.weight"].t())[..., :V]
    got_hid = out["hidden_states"][0].reshape(B, S, d)
    e_h = float((got_hid - hid).abs().max())
    e_l = float((out["logits"].reshape(B, S, V) - want_logits).abs().max())
    print(f"[oracle forward vs HF GPT2Model] hidden max abs diff {e_h:.2e}, logits {e_l:.2e}")
    assert float(hid.abs().max()) > 0.5 and float(want_logits.std()) > 1e-3 and torch.isfinite(want_logits).all()      # not a vacuous match
    assert e_h < 5e-5 and e_l < 5e-5
