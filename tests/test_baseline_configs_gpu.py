"""Parity at the BASELINE.json configurations themselves (GPU), not only at toy sizes:

  C2  interiornet-transformer generate(), 9 context views, batch 32 scenes (evaluate/evaluate_transformer.py:97-146)
      - the exact fp32 CUDA pipeline against the CPU oracle on the first 2 scenes (bit-exact codes, <= 1 LSB pixels);
      - the benchmarked mode (`mixed`, CUDA-graph replay) against the exact fp32 CUDA pipeline on all 32 scenes:
        encoder codes bit-exact (18 432 indices), generated codes / pixels within the stated tolerance.
  C1/C3  batch invariance of the encoder: the 4 golden images of the REAL reference embedded in a 288-image batch
      (exercises the wide conv's 32-bit index arithmetic at the benchmarked launch size).
  C5  full-size transformer, 19 context views: prefill_context + query == the full forward's last view, and both
      against the CPU oracle.
Tolerances: fp32 logits 5e-4 abs; bf16 transformer: argmax agreement >= 0.90, |dlogit| <= 6e-2; bf16 decoder: mean |d| <= 1.5 LSB.
"""
import os

import numpy as np
import pytest
import torch

from oracle import synth, vqgan_oracle as vo, migt_oracle as mo
from oracle.make_golden import vq_images
from viewformer_b200.config import VQGANConfig, MIGTConfig

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    """Full-size weights + the bench's own synthetic inputs (bench.synth_inputs, seed 1234)."""
    import bench
    vcfg, tcfg = VQGANConfig(), MIGTConfig(localization_weight="0")
    images, cams = bench.synth_inputs(32, 1234)
    vsd, tsd = synth.make_vqgan_state_dict(vcfg, 0), synth.make_migt_state_dict(tcfg, 0)
    return dict(vcfg=vcfg, tcfg=tcfg, images=images, cams=cams, vsd=vsd, tsd=tsd)


@pytest.fixture(scope="module")
def c2_exact(c2):
    """The exact fp32 CUDA pipeline on all 32 scenes."""
    from viewformer_b200 import VQGAN, MIGT, generate_batch_predictions
    cb = VQGAN(c2["vcfg"], precision="fp32").load_state_dict(c2["vsd"])
    tr = MIGT(c2["tcfg"], precision="fp32").load_state_dict(c2["tsd"])
    out = generate_batch_predictions(tr, cb, c2["images"], c2["cams"])
    codes = cb.encode_u8(c2["images"].cuda(), first_views=9)
    torch.cuda.synchronize()
    res = dict(gen_codes=out["generated_codes"].cpu(), gen_images=out["generated_images"].cpu(), codes=codes.cpu().reshape(32, 9, 8, 8))
    del cb, tr
    torch.cuda.empty_cache()
    return res


def test_c2_exact_pipeline_vs_oracle(c2, c2_exact):
    n = 2
    with torch.no_grad():
        torch.set_num_threads(min(32, os.cpu_count() or 1))
        want = mo.generate_batch_predictions(lambda d: mo.forward(c2["tsd"], c2["tcfg"], d, use_localization=False),
                                             lambda x: vo.encode(c2["vsd"], c2["vcfg"], x)[2],
                                             lambda c: vo.decode_code(c2["vsd"], c2["vcfg"], c), c2["tcfg"],
                                             c2["images"][:n], c2["cams"][:n], use_localization=False)
    mism = int((c2_exact["codes"][:n] != want["codes"][:, :9]).sum())
    print(f"[C2 fp32 vs oracle] encoder code mismatches {mism}/{n * 9 * 64}")
    assert mism == 0
    assert torch.equal(c2_exact["gen_codes"][:n], want["generated_codes"])
    d = (c2_exact["gen_images"][:n].int() - want["generated_images"].int()).abs()
    print(f"[C2 fp32 vs oracle] u8 pixel diff max {int(d.max())} nonzero {int((d > 0).sum())}/{d.numel()}")
    assert int(d.max()) <= 1


@pytest.mark.parametrize("precision", ["mixed"])
def test_c2_benchmarked_mode_vs_exact_on_all_scenes(c2, c2_exact, precision):
    from viewformer_b200 import VQGAN, MIGT, GraphedPredictions
    cb = VQGAN(c2["vcfg"], precision=precision).load_state_dict(c2["vsd"])
    tr = MIGT(c2["tcfg"], precision="bf16" if precision == "mixed" else precision).load_state_dict(c2["tsd"])
    gp = GraphedPredictions(tr, cb, 32, 10)
    out = gp(c2["images"].pin_memory(), c2["cams"].pin_memory())
    codes = cb.encode_u8(c2["images"].cuda(), first_views=9).cpu().reshape(32, 9, 8, 8)
    torch.cuda.synchronize()
    mism = int((codes != c2_exact["codes"]).sum())
    print(f"[C2 {precision} vs fp32] encoder code mismatches {mism}/18432")
    assert mism == 0                                                   # north_star: bit-exact token indices
    agree = float((out["generated_codes"].cpu() == c2_exact["gen_codes"]).float().mean())
    print(f"[C2 {precision} vs fp32] generated-code agreement {agree:.4f}")
    assert agree >= 0.90                                               # bf16 transformer: argmax of near-tied logits may flip
    # decoder tolerance on IDENTICAL codes
    px = cb.decode_code_u8(c2_exact["gen_codes"].cuda()).cpu().int()
    d = (px - c2_exact["gen_images"].int()).abs()
    print(f"[C2 {precision} vs fp32] decoder u8 diff on identical codes: max {int(d.max())} mean {float(d.float().mean()):.3f}")
    assert float(d.float().mean()) <= 1.5 and int(d.max()) <= 24


@pytest.mark.parametrize("precision", ["fp32", "mixed"])
def test_encoder_batch_invariance_288_images(golden_dir, precision):
    """The 4 golden images of the REAL reference placed at the front, the middle and the end of a 288-image batch."""
    from viewformer_b200 import VQGAN
    g = np.load(os.path.join(golden_dir, "vqgan_full.npz"))
    cfg = VQGANConfig()
    sd = synth.make_vqgan_state_dict(cfg, int(g["seed"]))
    model = VQGAN(cfg, precision=precision).load_state_dict(sd)
    x4 = vq_images(int(g["n_images"]), cfg.image_size, 1000 + int(g["seed"]))
    filler = vq_images(8, cfg.image_size, 4242)
    x = filler.repeat(36, 1, 1, 1).clone()                     # 288 images
    pos = [0, 1, 142, 287]
    for i, p in enumerate(pos):
        x[p] = x4[i]
    codes = model.encode(x)[2].cpu().numpy()
    for i, p in enumerate(pos):
        mism = int((codes[p] != g["codes"][i]).sum())
        assert mism == 0, f"image {i} at batch position {p}: {mism}/64 codes differ from the reference golden"
    # the repeated filler images encode identically wherever they sit in the batch
    ref8 = codes[8:16]
    for r in range(2, 35):
        assert (codes[8 * r:8 * r + 8] == ref8).all() or any(8 * r <= p < 8 * r + 8 for p in pos)


def test_c5_full_size_kv_cache_vs_full_forward_and_oracle():
    from viewformer_b200 import MIGT
    cfg = MIGTConfig(localization_weight="0")
    sd = synth.make_migt_state_dict(cfg, 5)
    B, T = 2, 20
    codes = synth.make_codes(B, T, seed=31)
    cams = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(B, T, seed=32))[0])
    ids = torch.cat([codes[:, :-1], torch.full_like(codes[:, :1], cfg.n_embeddings)], 1)
    with torch.no_grad():
        torch.set_num_threads(min(32, os.cpu_count() or 1))
        want = mo.forward(sd, cfg, dict(input_ids=ids[:1], poses=cams[:1]), use_localization=False)["logits"][:, -1]     # scene 0 only (CPU time)
    for precision, tol, agree_min in (("fp32", 5e-4, 1.0), ("bf16", 6e-2, 0.90)):
        model = MIGT(cfg, precision=precision).load_state_dict(sd)
        full = model(dict(input_ids=ids, poses=cams), last_only=True)["logits"][:, 0]
        cache = model.prefill_context(codes[:, :-1], cams[:, :-1].contiguous())
        q_codes, q_logits = model.query(cache, cams[:, -1].contiguous(), return_logits=True)
        torch.cuda.synchronize()
        e_cache = float((q_logits - full).abs().max())
        e_oracle = float((full[:1].cpu() - want).abs().max())
        agree = float((q_logits.argmax(-1) == full.argmax(-1)).float().mean())
        agree_o = float((full[:1].argmax(-1).cpu() == want.argmax(-1)).float().mean())
        print(f"[C5 {precision}] |query - full| max {e_cache:.3e}; |full - oracle| max {e_oracle:.3e}; argmax agreement query/full {agree:.3f}, full/oracle {agree_o:.3f}")
        assert e_cache < tol and e_oracle < tol
        assert agree >= agree_min and agree_o >= agree_min
        if precision == "fp32":
            assert torch.equal(q_codes.reshape(B, -1), full.argmax(-1).reshape(B, -1))
        del model, cache
        torch.cuda.empty_cache()
