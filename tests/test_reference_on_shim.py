"""The reference's OWN transformer sources executed here (container only): viewformer/models/migt.py, models/branching_attention.py,
models/utils.py (create_optimizer, WarmUp, AdamWeightDecay), utils/{tensorflow,geometry_tf,metrics,schedules}.py and
evaluate/evaluate_transformer{,_multictx}.py are loaded UNMODIFIED from /root/reference with oracle/tf_shim.py answering
``import tensorflow`` (TensorFlow 2.4.1 cannot be installed: no cp312 wheel, no network).  Every line of model wiring is the
reference's; the shim restates only leaf tensor ops, and its own semantics are checked below against TensorFlow's documented behaviour.

What this pins:
  * oracle/migt_oracle.py::forward == MIGT.call (migt.py:338-455) for the generate() call, the localisation call, compute_losses with
    label smoothing / pose multiplier / schedules / dynamic pose weights, explicit output_poses + localization_tokens, use_localization off;
  * the COMMITTED fixtures the GPU parity tests consume (tests/golden/migt_small.npz, migt_full.npz, migt_train_small.npz,
    migt_train_full.npz) are reproduced by the reference's code: forward, MIGT.train_step (GradientTape, clip, AdamWeightDecay under
    WarmUp(CosineDecay)) — so CUDA == fixture (tests -m gpu) and fixture == reference (here) close the chain;
  * generate_batch_predictions of both evaluation scripts == the oracle's restatements.
Skipped where /root/reference does not exist (the GPU box); tests/test_oracle_pinned.py::test_oracle_matches_reference_on_shim_fixture
checks the oracle against outputs of this setup committed as tests/golden/migt_reference_shim.npz everywhere."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import ref_loader, synth, migt_oracle as mo
from viewformer_b200.config import MIGTConfig, VQGANConfig

pytestmark = pytest.mark.skipif(not ref_loader.migt_available(), reason="reference sources not present (container-only test)")

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SMALL = dict(n_layer=2, n_head=4, d_model=64, sequence_size=4, n_embeddings=40, token_image_size=2, n_loss_skip=1)


@pytest.fixture(scope="module")
def tf():
    from oracle import tf_shim
    tf_shim.install()
    ref_loader.load_reference_migt()
    yield sys.modules["tensorflow"]
    tf_shim.uninstall()                       # later test modules must not see a `tensorflow` in sys.modules


def _var(model, key):
    suffix = "/" + key.replace(".", "/")
    hits = [v for v in model.variables if ("/" + v.name[:-2].replace(".", "/")).endswith(suffix)]
    assert len(hits) == 1, (key, [v.name for v in hits])
    return hits[0]


# ------------------------------------------------------------------------------------------------ the shim's own semantics
def test_shim_ops_follow_documented_tensorflow_semantics(tf):
    x = tf.reshape(tf.range(24), [2, 3, 4])
    assert x.shape.as_list() == [2, 3, 4] and tf.shape(x).tolist() == [2, 3, 4] and tf.rank(x) == 3
    a, b = tf.split(x, 2, axis=-1)                                   # an int = NUMBER of equal parts (torch.split: size of a part)
    assert a.shape == (2, 3, 2) and b[0, 0].tolist() == [2, 3]
    p, q = tf.split(x, (1, 3), axis=-1)
    assert p.shape == (2, 3, 1) and q.shape == (2, 3, 3)
    assert tf.repeat(tf.range(3), 2).tolist() == [0, 0, 1, 1, 2, 2]  # element-wise, not tiling
    assert tf.tile(tf.range(3), [2]).tolist() == [0, 1, 2, 0, 1, 2]
    assert tf.one_hot([0, 2], 3).tolist() == [[1, 0, 0], [0, 0, 1]]
    assert tf.gather(tf.constant([[1., 2.], [3., 4.], [5., 6.]]), tf.constant([[2, 0]])).tolist() == [[[5, 6], [1, 2]]]
    assert tf.constant(3).dtype == tf.int32 and tf.constant(3.0).dtype == tf.float32 and tf.constant(3, "float32").dtype == tf.float32
    assert tf.constant(np.zeros(2, np.uint8)).dtype == tf.uint8       # numpy arrays keep their dtype
    g = tf.constant([3.0, 4.0])
    assert torch.allclose(tf.clip_by_norm(g, 1.0), torch.tensor([0.6, 0.8])) and torch.equal(tf.clip_by_norm(g, 10.0), torch.tensor([3.0, 4.0]))
    assert torch.allclose(tf.linalg.l2_normalize(g, axis=-1), torch.tensor([0.6, 0.8]))
    assert torch.equal(tf.linalg.l2_normalize(tf.zeros([2]), axis=-1, epsilon=1e-12), torch.zeros(2))      # x * rsqrt(max(sum x^2, eps))
    u8 = tf.constant(np.array([0, 1, 128, 255], np.uint8))
    f = tf.image.convert_image_dtype(u8, tf.float32)
    assert torch.equal(f, u8.float() * torch.tensor(1 / 255, dtype=torch.float32))
    assert tf.image.convert_image_dtype(tf.constant([0.0, 0.5, 0.999, 1.0]), tf.uint8).tolist() == [0, 127, 255, 255]    # x * 255.5, truncated
    logits = tf.constant([[1.0, 2.0, 3.0]])
    lp = torch.log_softmax(logits, -1)
    assert torch.allclose(tf.nn.sparse_softmax_cross_entropy_with_logits(tf.constant([2]), logits), -lp[:, 2])
    assert torch.allclose(tf.nn.softmax_cross_entropy_with_logits(tf.constant([[0.5, 0.5, 0.0]]), logits), -(0.5 * lp[:, 0] + 0.5 * lp[:, 1]))
    assert torch.allclose(tf.losses.mse(tf.constant([[1.0, 3.0]]), tf.constant([[0.0, 0.0]])), torch.tensor([5.0]))
    assert torch.allclose(tf.nn.gelu(tf.constant([1.0])), torch.tensor([0.8413447]))                       # exact erf form
    assert torch.equal(tf.matmul(tf.ones([2, 3]), tf.ones([4, 3]), transpose_b=True), torch.full((2, 4), 3.0))
    assert tf.ones(tf.shape(tf.zeros([5, 2]))[0], "float32").shape == (5,)                                  # scalar shape -> 1-D
    init = tf.keras.initializers.TruncatedNormal(0.02)                # positional argument = MEAN (stddev stays 0.05), as in Keras
    w = init([4000])
    assert abs(float(w.mean()) - 0.02) < 5e-3 and float((w - 0.02).abs().max()) <= 0.1 + 1e-6
    ln = tf.keras.layers.LayerNormalization(epsilon=1e-5, name="ln")
    y = ln(tf.constant([[1.0, 2.0, 3.0, 6.0]]))
    assert abs(float(y.mean())) < 1e-6 and [v.name for v in ln.variables] == ["ln/gamma:0", "ln/beta:0"]
    sched = tf.keras.experimental.CosineDecay(2.0, 10)
    assert abs(float(sched(0)) - 2.0) < 1e-6 and abs(float(sched(5)) - 1.0) < 1e-6 and abs(float(sched(20))) < 1e-6


def test_shim_adam_matches_torch_adam(tf):
    """The Adam base under the reference's AdamWeightDecay: Keras' epsilon-outside-the-bias-correction form.  Against torch.optim.Adam,
    whose epsilon sits inside (eps_hat = eps * sqrt(1 - b2^t)), the two agree when eps is negligible."""
    w = tf.Variable(torch.linspace(-1, 1, 7))
    opt = tf.keras.optimizers.Adam(learning_rate=0.01, beta_1=0.9, beta_2=0.999, epsilon=1e-12)
    p = torch.nn.Parameter(torch.linspace(-1, 1, 7))
    ref = torch.optim.Adam([p], lr=0.01, betas=(0.9, 0.999), eps=1e-12)
    for step in range(4):
        g = torch.sin(torch.arange(7.0) + step)
        opt.apply_gradients([(g, w)])
        p.grad = g.clone()
        ref.step()
    assert float((w.detach() - p.detach()).abs().max()) < 1e-6 and int(opt.iterations) == 4


# ------------------------------------------------------------------------------------------------ MIGT.call == oracle.forward
def _inputs(cfg, variant, B=2, T=4):
    codes = synth.make_codes(B, T, n_embed=cfg.n_embeddings, side=cfg.token_image_size, seed=5)
    cams = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(B, T, seed=6))[0])
    if variant == "generate":           # last view masked (evaluate_transformer.py:120-123)
        return dict(input_ids=torch.cat([codes[:, :-1], torch.full_like(codes[:, :1], cfg.n_embeddings)], 1), poses=cams), False
    if variant == "localize":           # one pose fewer than views (evaluate_transformer.py:135-137)
        return dict(input_ids=codes, poses=cams[:, :-1]), False
    if variant == "losses":
        return dict(input_ids=codes, poses=cams), True
    if variant == "explicit":           # the multictx call (evaluate_transformer_multictx.py:61-72)
        return dict(input_ids=codes, poses=cams, output_poses=cams.flip(1).contiguous(), localization_tokens=codes.flip(0).contiguous()), False
    raise ValueError(variant)


CASES = [("generate", {}), ("localize", {}), ("losses", {}), ("explicit", {}),
         ("generate", dict(localization_weight="0")), ("losses", dict(localization_weight="0")),
         ("losses", dict(label_smoothing=0.1, pose_multiplier=0.3, image_generation_weight=0.7, localization_weight="cosine(2.0,0.5,100)")),
         ("losses", dict(use_dynamic_pose_loss=True, localization_weight="warmup(linear(1.0,0.2,50),10)")),
         ("generate", dict(token_image_size=4, n_head=2, sequence_size=6))]


@pytest.mark.parametrize("variant,extra", CASES)
def test_reference_call_equals_oracle_forward(tf, variant, extra):
    kw = dict(SMALL, **extra)
    cfg = MIGTConfig(**kw)
    sd = synth.make_migt_state_dict(cfg, 3)
    dyn = [0.3, -2.0] if kw.get("use_dynamic_pose_loss") else None
    model = ref_loader.build_reference_migt(sd, dynamic_pose_weights=dyn, **kw)
    if dyn is not None:
        sd = dict(sd, **{"pose_loss_weighting_criterion.pos_ori_weights": torch.tensor(dyn)})
    inputs, losses = _inputs(cfg, variant, T=cfg.sequence_size)
    step = 7
    model._train_counter.assign(step)
    with torch.no_grad():
        r = model({k: v.clone() for k, v in inputs.items()}, compute_losses=losses, training=False)
        lw = float(model.localization_weight(step)) if model.use_localization else 0.0
        o = mo.forward(sd, cfg, inputs, compute_losses=losses, use_localization=model.use_localization, localization_weight=lw)
    assert len(r["hidden_states"]) == len(o["hidden_states"])
    keys = ["logits", "loss"] + [k for k in ("pose_prediction", "ce_loss", "pose_loss", "pose_pos_loss", "pose_ori_loss") if k in r]
    assert all(k in o for k in keys) and ("pose_prediction" in r) == ("pose_prediction" in o)
    for k in keys:
        a, b = torch.as_tensor(r[k]).float(), torch.as_tensor(o[k]).float()
        assert a.shape == b.shape and torch.isfinite(a).all(), k
        assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(a.abs().max())), (k, float((a - b).abs().max()))
    for a, b in zip(r["hidden_states"], o["hidden_states"]):
        assert float((a - b).abs().max()) <= 1e-6 * max(1.0, float(a.abs().max()))
    if losses and model.use_localization:
        assert abs(float(r["localization_weight"]) - lw) < 1e-7


# ------------------------------------------------------------------------------------------------ committed fixtures == reference
@pytest.mark.parametrize("tag,overrides,B,T", [("small", dict(n_layer=2, n_head=4, d_model=128, sequence_size=4, n_loss_skip=1), 2, 4), ("full", {}, 1, 10)])
def test_reference_reproduces_committed_forward_fixture(tf, tag, overrides, B, T):
    """tests/golden/migt_{small,full}.npz (what tests/test_models_gpu.py compares the CUDA transformer with; `full` = the 12-layer,
    d = 768 BASELINE model) recomputed by the reference's MIGT.call + reduce_cameras."""
    G = np.load(os.path.join(GOLDEN, f"migt_{tag}.npz"))
    cfg = MIGTConfig(**overrides)
    sd = synth.make_migt_state_dict(cfg, 3)
    model = ref_loader.build_reference_migt(sd, **overrides)
    codes = synth.make_codes(B, T, seed=5)
    cams = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(B, T, seed=6))[0])
    ids = torch.cat([codes[:, :-1], torch.full_like(codes[:, :1], cfg.n_embeddings)], 1)
    with torch.no_grad():
        last = model(dict(input_ids=ids, poses=cams), training=False)["logits"][:, -1]
        o2 = model(dict(input_ids=codes, poses=cams[:, :-1]), training=False)
        pose = model.reduce_cameras(o2["pose_prediction"][:, -1:], -2)
    assert float((last[:1] - torch.from_numpy(G["logits_last"])).abs().max()) < 2e-5
    assert np.array_equal(last.argmax(-1).numpy(), G["argmax_last"])
    assert float((pose - torch.from_numpy(G["pose_last"])).abs().max()) < 1e-6


def test_reference_train_step_reproduces_committed_train_fixture(tf):
    """Three calls of the reference's MIGT.train_step — GradientTape over its forward, per-tensor clip, its own create_optimizer →
    AdamWeightDecay under WarmUp(CosineDecay) — land on the post-step weights of tests/golden/migt_train_small.npz, the file the CUDA
    trainer is compared with in tests/test_train_gpu.py."""
    from oracle import make_golden as mg
    kw = dict(mg.MIGT_TRAIN)
    cfg = MIGTConfig(**kw)
    sd = synth.make_migt_state_dict(cfg, 9)
    model = ref_loader.build_reference_migt(sd, **kw)
    utils = sys.modules["viewformer.models.utils"]
    opt, _ = utils.create_optimizer(cfg.learning_rate, num_train_steps=cfg.total_steps, num_warmup_steps=mg.MIGT_TRAIN_WARMUP,
                                    weight_decay_rate=cfg.weight_decay)
    assert type(opt).__name__ == "AdamWeightDecay"
    model.compile(optimizer=opt)
    G = np.load(os.path.join(GOLDEN, "migt_train_small.npz"))
    names = [str(n) for n in G["names"]]
    gen = torch.Generator().manual_seed(77)
    probe = {k: torch.randn(sd[k].shape, generator=gen) for k in names}
    keep = [k[3:] for k in G.files if k.startswith("p0.")]
    assert len(keep) >= 6
    losses = []
    for step in range(3):
        codes = synth.make_codes(2, 4, n_embed=cfg.n_embeddings, seed=50 + step)
        cams = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(2, 4, seed=60 + step))[0])
        model._train_counter.assign(step)
        metrics = model.train_step((cams, codes))
        losses.append(float(metrics["loss"]))                        # running mean of the per-step losses (tf.metrics.Mean)
        pdot = np.array([float((_var(model, k).detach() * probe[k]).sum()) for k in names])
        assert np.max(np.abs(pdot - G[f"pdot{step}"]) / (np.abs(G[f"pdot{step}"]) + 1e-3)) < 2e-5, step
        for k in keep:
            assert float((_var(model, k).detach() - torch.from_numpy(G[f"p{step}.{k}"])).abs().max()) < 1e-6, (step, k)
    want = np.cumsum([float(G[f"loss{i}"]) for i in range(3)]) / np.arange(1, 4)
    assert np.allclose(losses, want, rtol=1e-5)
    assert int(opt.iterations) == 3


class _RecordingOptimizer:
    """Captures what MIGT.train_step hands to apply_gradients (after its clipping)."""

    def apply_gradients(self, grads_and_vars):
        self.pairs = [(g, v) for g, v in grads_and_vars]


def test_reference_train_step_gradients_full_size(tf):
    """The 12-layer, d = 768 model (MIGTConfig defaults), label smoothing 0.1, localization weight 0.7: loss terms and the gradient of
    every tensor as the reference's train_step computes them == tests/golden/migt_train_full.npz (norm and a random projection)."""
    kw = dict(dropout=0.0, label_smoothing=0.1, localization_weight="0.7", total_steps=100, learning_rate=1e-4)
    cfg = MIGTConfig(**kw)
    sd = synth.make_migt_state_dict(cfg, 13)
    model = ref_loader.build_reference_migt(sd, **kw)
    rec = _RecordingOptimizer()
    model.compile(optimizer=rec)
    codes = synth.make_codes(1, 5, n_embed=cfg.n_embeddings, seed=70)
    cams = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(1, 5, seed=71))[0])
    metrics = model.train_step((cams, codes))
    G = np.load(os.path.join(GOLDEN, "migt_train_full.npz"))
    names = [str(n) for n in G["names"]]
    assert abs(float(metrics["loss"]) - float(G["loss"])) < 1e-5 * abs(float(G["loss"]))
    assert abs(float(metrics["ce_loss"]) - float(G["ce"].mean())) < 1e-5 and abs(float(metrics["pose_loss"]) - float(G["pose"].mean())) < 1e-5
    grads = {id(v): g for g, v in rec.pairs}
    gen = torch.Generator().manual_seed(78)
    probe = {k: torch.randn(sd[k].shape, generator=gen) for k in names}
    for i, k in enumerate(names):
        g = grads[id(_var(model, k))]
        g = torch.zeros_like(sd[k]) if g is None else g
        assert abs(float(g.norm()) - G["gnorm"][i]) <= 2e-4 * G["gnorm"][i] + 1e-7, k
        assert abs(float((g * probe[k]).sum()) - G["gdot"][i]) <= 2e-4 * G["gnorm"][i] * float(probe[k].norm()) + 1e-7, k


def test_reference_dynamic_pose_weights_gradient(tf):
    """use_dynamic_pose_loss: the learned (pos, ori) log-weights are trainable variables of the reference model; their gradient through
    DynamicLossWeightingCriterion (migt.py:107-120) == autograd through the oracle."""
    kw = dict(SMALL, dropout=0.0, use_dynamic_pose_loss=True, localization_weight="0.5")
    cfg = MIGTConfig(**kw)
    sd = synth.make_migt_state_dict(cfg, 4)
    model = ref_loader.build_reference_migt(sd, dynamic_pose_weights=[0.2, -1.5], **kw)
    rec = _RecordingOptimizer()
    model.compile(optimizer=rec)
    inputs, _ = _inputs(cfg, "losses")
    model.train_step((inputs["poses"], inputs["input_ids"]))
    g_ref = [g for g, v in rec.pairs if v.name.startswith("pos_ori_weights")]
    assert len(g_ref) == 1
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    leaves["pose_loss_weighting_criterion.pos_ori_weights"] = torch.tensor([0.2, -1.5], requires_grad=True)
    o = mo.forward(leaves, cfg, inputs, compute_losses=True, localization_weight=0.5)
    o["loss"].mean().backward()
    assert float((g_ref[0] - leaves["pose_loss_weighting_criterion.pos_ori_weights"].grad).abs().max()) < 1e-6
    gw = [g for g, v in rec.pairs if v.name.endswith("h.0/attn/c_attn/weight:0")][0]
    assert float((gw - leaves["h.0.attn.c_attn.weight"].grad).abs().max()) < 1e-6


# ------------------------------------------------------------------------------------------------ the evaluation callers
@pytest.mark.parametrize("augment,loc", [("relative", "1"), ("relative", "0"), ("no", "1")])
def test_reference_generate_batch_predictions_equal_oracle(tf, augment, loc):
    """evaluate_transformer.py:97-146 and evaluate_transformer_multictx.py:37-95 run as shipped — with the REAL reference torch codebook
    behind an NHWC adapter and the real MIGT on the shim — against oracle.generate_batch_predictions{,_multictx}: uint8 views, codes and
    cameras identical."""
    from oracle import make_golden as mg
    ev, mc = ref_loader.load_reference_evaluate()
    vcfg = VQGANConfig(**mg.SMALL_VQ)
    vq = ref_loader.build_reference_vqgan(synth.make_vqgan_state_dict(vcfg, 0), **mg.SMALL_VQ)
    kw = dict(n_layer=2, n_head=4, d_model=64, sequence_size=4, n_embeddings=vcfg.n_embed, token_image_size=8, n_loss_skip=1,
              augment_poses=augment, localization_weight=loc)
    cfg = MIGTConfig(**kw)
    sd = synth.make_migt_state_dict(cfg, 3)
    model = ref_loader.build_reference_migt(sd, **kw)
    images = synth.make_images_uint8(2, 4, size=vcfg.image_size, seed=11)
    cams = synth.make_cameras(2, 4, seed=12)
    codebook = ref_loader.ReferenceCodebookNHWC(vq)
    use_loc = model.use_localization

    def fwd(inp):
        return mo.forward(sd, cfg, inp, use_localization=use_loc)

    def enc(x):
        return vq.encode(x)[-1]
    with torch.no_grad():
        r = ev.generate_batch_predictions(model, codebook, images.clone(), cams.clone())
        o = mo.generate_batch_predictions(fwd, enc, vq.decode_code, cfg, images, cams, use_localization=use_loc)
    assert r["generated_images"].dtype == torch.uint8 and tuple(r["generated_images"].shape) == (2, 32, 32, 3)
    assert float(r["generated_images"].float().std()) > 5
    for k in ("ground_truth_images", "generated_images"):
        assert torch.equal(torch.as_tensor(r[k]).as_subclass(torch.Tensor), o[k]), k
    for k in ("ground_truth_cameras", "generated_cameras"):
        assert float((r[k] - o[k]).abs().max()) < 1e-6, k
    if use_loc:                                                       # the 3-stream script needs the pose head
        with torch.no_grad():
            r = mc.generate_batch_predictions(model, codebook, images.clone(), cams.clone())
            o = mo.generate_batch_predictions_multictx(fwd, enc, vq.decode_code, cfg, images, cams)
        assert tuple(r["generated_images"].shape) == (2, 4, 32, 32, 3)
        assert torch.equal(torch.as_tensor(r["generated_images"]).as_subclass(torch.Tensor), o["generated_images"])
        assert float((r["generated_cameras"] - o["generated_cameras"]).abs().max()) < 1e-6


def test_reference_allimg_script_equals_oracle(tf):
    """evaluate_transformer_multictx_allimg.py:15-88 as shipped (encode_images, transformer_predict, run_with_batchsize, decode_code):
    what viewformer_b200.evaluate mirrors and tests/test_eval_gpu.py::test_transformer_predict_and_steps_match_oracle compares with
    oracle.generate_batch_predictions_multictx.  The script blanks the LAST context camera with zeros where the multictx script keeps
    it (allimg:28 vs multictx:64) — the outputs must not depend on it (the query streams only see strictly earlier views)."""
    from oracle import make_golden as mg
    ref_loader.load_reference_evaluate()
    R = os.path.join(ref_loader.REFERENCE_ROOT, "viewformer", "evaluate", "evaluate_transformer_multictx_allimg.py")
    al = ref_loader._load("viewformer.evaluate.evaluate_transformer_multictx_allimg", R)
    vcfg = VQGANConfig(**mg.SMALL_VQ)
    vq = ref_loader.build_reference_vqgan(synth.make_vqgan_state_dict(vcfg, 0), **mg.SMALL_VQ)
    kw = dict(n_layer=2, n_head=4, d_model=64, sequence_size=4, n_embeddings=vcfg.n_embed, token_image_size=8, n_loss_skip=1)
    cfg = MIGTConfig(**kw)
    sd = synth.make_migt_state_dict(cfg, 3)
    model = ref_loader.build_reference_migt(sd, **kw)
    codebook = ref_loader.ReferenceCodebookNHWC(vq)
    images = synth.make_images_uint8(3, 4, size=vcfg.image_size, seed=41)
    cams = synth.make_cameras(3, 4, seed=42)
    with torch.no_grad():
        codes = al.encode_images(images.clone(), codebook_model=codebook)
        gen_cams, gen_codes = al.run_with_batchsize(al.transformer_predict, 2, cams.clone(), codes, transformer_model=model)
        gen_images = al.decode_code(gen_codes, codebook_model=codebook)
        o = mo.generate_batch_predictions_multictx(lambda d: mo.forward(sd, cfg, d), lambda x: vq.encode(x)[-1], vq.decode_code, cfg, images, cams)
        want_codes = vq.encode(mo.images_to_float(images.reshape(-1, *images.shape[2:])).permute(0, 3, 1, 2).contiguous())[-1].reshape(3, 4, 8, 8)
    assert tuple(codes.shape) == (3, 4, 8, 8) and torch.equal(torch.as_tensor(codes).long(), want_codes.long())
    assert torch.equal(torch.as_tensor(gen_codes).long(), o["generated_codes"].long())
    assert float((torch.as_tensor(gen_cams) - o["generated_cameras"]).abs().max()) < 1e-6
    assert torch.equal(torch.as_tensor(gen_images).as_subclass(torch.Tensor), o["generated_images"])


def test_process_batch_equals_the_reference_function(tf):
    """train/train_transformer.py:31-64 (the pose augmentation mapped over every training sample): the function's own source text,
    cut out of the reference file with ast (the script around it needs the whole Keras training stack to import) and executed over the
    shim with the reference's geometry_tf, against viewformer_b200.data.process_batch — every augment mode, both splits, random modes
    under the same seed (the shim draws from torch's global generator, in the order the reference's expressions are evaluated)."""
    import ast
    import math
    from viewformer_b200.data import process_batch
    ref_loader.load_reference_migt()
    geometry = sys.modules["viewformer.utils.geometry_tf"]
    path = os.path.join(ref_loader.REFERENCE_ROOT, "viewformer", "train", "train_transformer.py")
    tree = ast.parse(open(path).read())
    fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "process_batch"]
    assert len(fn) == 1
    ns = dict(tf=tf, geometry=geometry, math=math)
    exec(compile(ast.Module(body=fn, type_ignores=[]), path, "exec"), ns)
    ref_fn = ns["process_batch"]
    cams = synth.make_cameras(3, 5, seed=8)                             # [B, T, 7]: the loader maps over [T, 7] windows, batches also work
    toks = synth.make_codes(3, 5, n_embed=16, side=2, seed=9)
    for x in (cams[0], cams):
        for augment in ("relative", "no", "simple", "advanced"):
            for split in ("train", "test"):
                torch.manual_seed(123)
                want, wt = ref_fn(tf.constant(x.numpy()), toks, augment, split)
                torch.manual_seed(123)
                got, gt = process_batch(x.clone(), toks, augment, split)
                assert gt is toks and wt is toks
                assert float((torch.as_tensor(want) - got).abs().max()) < 2e-6, (augment, split)
                assert bool((got[..., 3] >= 0).all()) and float((got[..., 3:].norm(dim=-1) - 1).abs().max()) < 1e-5
                if augment in ("simple", "advanced"):                    # random modes act on the train split only
                    same = float((got - process_batch(x.clone(), toks, "no", split)[0]).abs().max()) < 1e-6
                    assert same == (split != "train")
    with pytest.raises(ValueError):
        process_batch(cams[0], toks, "bogus", "train")


# ---------------------------------------------------------------------------------------- evaluation metrics (utils/metrics.py, Evaluator)
def test_shim_image_ops_follow_documented_tensorflow_semantics(tf):
    """The leaf ops utils/metrics.py adds to the shim's surface: tf.nn.depthwise_conv2d (NHWC, filter [fh,fw,in,mult], output channel
    k*mult + q), tf.image.psnr, tf.image.convert_image_dtype(uint8 -> float = /255), Keras MeanSquaredError (CAST, no rescale)."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((2, 5, 6, 3)).astype(np.float32)
    w = rng.standard_normal((2, 3, 3, 2)).astype(np.float32)
    y = tf.nn.depthwise_conv2d(tf.constant(x), tf.constant(w), strides=[1, 1, 1, 1], padding="VALID").numpy()
    want = np.zeros((2, 4, 4, 6), np.float32)
    for k in range(3):
        for q in range(2):
            for i in range(4):
                for j in range(4):
                    want[:, i, j, k * 2 + q] = (x[:, i:i + 2, j:j + 3, k] * w[:, :, k, q]).sum((1, 2))
    assert y.shape == want.shape and np.allclose(y, want, atol=1e-5)
    a = rng.integers(0, 256, (2, 8, 8, 3), dtype=np.uint8)
    b = rng.integers(0, 256, (2, 8, 8, 3), dtype=np.uint8)
    fa, fb = tf.image.convert_image_dtype(a, "float32"), tf.image.convert_image_dtype(b, "float32")
    assert fa.dtype == tf.float32 and np.allclose(fa.numpy(), a / 255.0, atol=1e-7)
    mse = ((a / 255.0 - b / 255.0) ** 2).mean((1, 2, 3))
    assert np.allclose(tf.image.psnr(fa, fb, 1).numpy(), 10 * np.log10(1 / mse), atol=1e-4)
    m = tf.keras.metrics.MeanSquaredError("mse")
    m.update_state(a, b)
    assert abs(float(m.result()) - ((a.astype(np.float64) - b) ** 2).mean()) < 1e-2               # 0..255 scale: uint8 is cast, not rescaled
    m = tf.keras.metrics.MeanAbsoluteError("mae")
    m.update_state(a, b)
    assert abs(float(m.result()) - np.abs(a.astype(np.float64) - b).mean()) < 1e-3


def _reference_evaluator():
    ev, _ = ref_loader.load_reference_evaluate()
    metrics = sys.modules["viewformer.utils.metrics"]
    metrics.LPIPSMetric._lpips_pool["vgg"] = lambda a, b: torch.zeros(a.shape[0])       # LPIPS needs VGG weights: not on this path
    return ev, metrics


def test_evaluator_fixture_is_what_the_reference_evaluator_computes(tf, golden_dir):
    """tests/golden/evaluator_reference_shim.npz (what viewformer_b200.metrics.Evaluator is held to on the GPU,
    tests/test_vs_reference_evaluator_gpu.py) is reproduced by the reference's own Evaluator (evaluate_transformer.py:22-67)."""
    from oracle import make_golden as G
    ev, _ = _reference_evaluator()
    g = np.load(os.path.join(golden_dir, "evaluator_reference_shim.npz"))
    for tag, n, gs, ns, image_size, seed in G.EVALUATOR_CASES:
        gt, gen = synth.make_metric_pair(n, gs, ns, seed)
        assert [int(gt.sum()), int(gen.sum())] == g[f"{tag}.input_sums"].tolist()
        E = ev.Evaluator(image_size)
        E.update_with_image(tf.convert_to_tensor(gt.numpy()), tf.convert_to_tensor(gen.numpy()))
        r = E.result()
        for k in ("mse", "rmse", "mae", "psnr", "ssim"):
            assert abs(r[k] - float(g[f"{tag}.{k}"])) <= 1e-6 * abs(r[k]), (tag, k)
    # what follows from HOW the reference calls its metrics: mse / mae on the 0..255 scale, rmse^2 == mse
    assert abs(float(g["same.mse"]) - float(g["same.rmse"]) ** 2) < 1e-3 * float(g["same.mse"]) and float(g["same.mae"]) > 1.0


def test_ssim_of_the_gpu_tests_and_the_k1_quirk_follow_the_reference(tf):
    """(a) the fp64 restatement tests/test_eval_gpu.py::ssim_ref that the CUDA kernel is compared with equals the reference's ssim()
    (utils/metrics.py:17-73) with its default K1 = 0.01; (b) SSIMMetric — the class the evaluators use — passes 1 as the THIRD positional
    argument, which is K1, not the data range (metrics.py:183): its value is ssim(K1 = 1), measurably different."""
    import importlib
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    T = importlib.import_module("test_eval_gpu")
    _, metrics = _reference_evaluator()
    gt, gen = synth.make_metric_pair(4, 64, 64, 5)
    A, B = gt.float() / 255, gen.float() / 255
    ref_default = torch.as_tensor(metrics.ssim(A.numpy(), B.numpy())).double()
    assert torch.allclose(T.ssim_ref(A, B), ref_default, atol=2e-6)
    m = metrics.SSIMMetric()
    m.update_state(gt.numpy(), gen.numpy())
    ref_k1 = torch.as_tensor(metrics.ssim(A.numpy(), B.numpy(), 1)).double().mean()
    assert abs(float(m.result()) - float(ref_k1)) < 1e-6
    assert abs(float(ref_k1) - float(ref_default.mean())) > 1e-4


def test_allow_nan_mean_counts_nan_as_zero_with_full_weight_like_the_reference(tf):
    """metrics.py:76-88 replaces NaN by 0 and then computes the weight from is_nan of the CLEANED values: a NaN camera error counts as 0
    with weight 1.  viewformer_b200.metrics.Mean(nan="zero") reproduces that for loc-angle / loc-dist."""
    from viewformer_b200.metrics import Evaluator
    _, metrics = _reference_evaluator()
    a = torch.tensor([[0., 0, 0, 1, 0, 0, 0], [1, 1, 1, 1, 0, 0, 0], [2, 0, 0, 1, 0, 0, 0]])
    b = torch.tensor([[3., 4, 0, 1, 0, 0, 0], [float("nan"), 1, 1, 1, 0, 0, 0], [2, 0, 1, 1, 0, 0, 0]])
    m = metrics.CameraPositionError()
    m.update_state(tf.convert_to_tensor(a.numpy()), tf.convert_to_tensor(b.numpy()))
    assert abs(float(m.result()) - 2.0) < 1e-6                                           # (5 + 0 + 1) / 3, not (5 + 1) / 2
    ours = Evaluator()
    ours.update_with_camera(a, b)
    assert abs(ours.result()["loc-dist"] - 2.0) < 1e-6


def test_multi_context_evaluator_equals_the_reference_class(tf):
    """evaluate_transformer_multictx.py:13-34 as shipped against viewformer_b200.metrics.MultiContextEvaluator: same ctxNN keys, same
    camera statistics per context size (the image half of each Evaluator needs the GPU: tests/test_vs_reference_evaluator_gpu.py)."""
    from viewformer_b200.metrics import MultiContextEvaluator
    _reference_evaluator()
    _, mc = ref_loader.load_reference_evaluate()
    B, T = 4, 4
    gt_cams = synth.make_cameras(1, B, seed=31)[0]                                     # [B, 7] the target view's camera
    g = torch.Generator().manual_seed(32)
    gen_cams = gt_cams[:, None].repeat(1, T, 1) + 0.2 * torch.randn((B, T, 7), generator=g)
    gt_img, _ = synth.make_metric_pair(B, 16, 16, 33)
    gen_img = torch.stack([synth.make_metric_pair(B, 16, 16, 40 + i)[1] for i in range(T)], 1)
    ref = mc.MultiContextEvaluator(T)
    ref.update_state(tf.convert_to_tensor(gt_cams.numpy()), tf.convert_to_tensor(gen_cams.numpy()), tf.convert_to_tensor(gt_img.numpy()),
                     tf.convert_to_tensor(gen_img.numpy()))
    ours = MultiContextEvaluator(T)
    ours.update_state(gt_cams, gen_cams)                                               # cameras only: nothing touches the device
    r, o = ref.result(), ours.result()
    assert list(r) == list(o) == ["ctx01", "ctx02", "ctx03"]
    for k in r:
        for m in ("loc-angle", "loc-dist", "loc-angle-med", "loc-dist-med"):
            assert abs(float(r[k][m]) - o[k][m]) < 2e-6 * max(1.0, abs(o[k][m])), (k, m)
    assert set(ours.get_progress_bar_info()) == {"img_psnr", "cam_loc", "cam_ang"}
    assert abs(float(ref.get_progress_bar_info()["cam_loc"]) - ours.get_progress_bar_info()["cam_loc"]) < 2e-6


def test_checkpoint_object_paths_are_the_reference_models_attribute_names(tf):
    """A TF2 object-based checkpoint (what model.save_weights writes and utils/tensorflow.py:57-61 loads) is keyed by the Python
    attribute names along the path from the model to each variable.  Walks the REAL reference MIGT object along
    tf_checkpoint.object_paths(key)[0] for every state_dict key and requires the variable found there to be the one the key names
    (found independently by its Keras variable name) — this is what caught 'wpe' (a variable hanging off the model itself, not
    wpe/embeddings) and pose_criterion/pose_classifier (not pose_classifier at the root)."""
    from viewformer_b200 import tf_checkpoint as tfc
    kw = dict(SMALL, use_dynamic_pose_loss=True)
    cfg = MIGTConfig(**kw)
    sd = synth.make_migt_state_dict(cfg, 3)
    model = ref_loader.build_reference_migt(sd, dynamic_pose_weights=[0.3, -2.0], **kw)
    obj = model                                                        # the dynamic pose-loss weights travel outside synth's state dict
    for part in tfc.object_paths("pose_loss_weighting_criterion.pos_ori_weights")[0].split("/"):
        obj = getattr(obj, part)
    assert tuple(obj.shape) == (2,) and torch.as_tensor(obj).detach().tolist() == pytest.approx([0.3, -2.0])
    for k, v in sd.items():
        obj = model
        for part in tfc.object_paths(k)[0].split("/"):
            obj = obj[int(part)] if part.isdigit() else getattr(obj, part)
        assert tuple(obj.shape) == tuple(v.shape), k
        assert obj is _var(model, k), k                                # same variable object as the one found by name
        assert float((torch.as_tensor(obj).detach() - v).abs().max()) == 0.0
    # the literal dotted form is NOT an attribute path of the real model for these two groups
    assert not hasattr(model, "pose_classifier") and not hasattr(model.wpe, "embeddings")


def test_codebook_evaluation_script_equals_oracle_and_fixture(tf, golden_dir):
    """evaluate/evaluate_codebook.py as shipped: its generate_batch_predictions (encode -> decode round trip, BASELINE configs[0]) with the
    REAL torch codebook equals the oracle's encode / decode_code / float_to_images chain — what viewformer_b200.evaluate
    .generate_codebook_predictions is compared with on the GPU — and its Evaluator reports the image numbers of the committed fixture."""
    from oracle import make_golden as mg
    ref_loader.load_reference_evaluate()
    _reference_evaluator()
    path = os.path.join(ref_loader.REFERENCE_ROOT, "viewformer", "evaluate", "evaluate_codebook.py")
    ec = ref_loader._load("viewformer.evaluate.evaluate_codebook", path)
    vcfg = VQGANConfig(**mg.SMALL_VQ)
    vsd = synth.make_vqgan_state_dict(vcfg, 0)
    vq = ref_loader.build_reference_vqgan(vsd, **mg.SMALL_VQ)
    for size in (vcfg.image_size, 48):                                  # 48 -> 32: the dataset resize rule in front of the encoder
        images = synth.make_images_uint8(1, 3, size=size, seed=51)[0]
        with torch.no_grad():
            r = ec.generate_batch_predictions(ref_loader.ReferenceCodebookNHWC(vq), tf.convert_to_tensor(images.numpy()))
            x = images if size == vcfg.image_size else torch.from_numpy(sys.modules["viewformer.data._common"].resize(images.numpy(), vcfg.image_size))
            codes = vq.encode(mo.images_to_float(x).permute(0, 3, 1, 2).contiguous())[-1]
            want = mo.float_to_images(vq.decode_code(codes).permute(0, 2, 3, 1))
        assert torch.equal(torch.as_tensor(r["generated_images"]).as_subclass(torch.Tensor), want)
        assert torch.equal(torch.as_tensor(r["ground_truth_images"]).as_subclass(torch.Tensor), images)
    g = np.load(os.path.join(golden_dir, "evaluator_reference_shim.npz"))
    gt, gen = synth.make_metric_pair(5, 64, 64, 21)
    E = ec.Evaluator()
    E.update_state(tf.convert_to_tensor(gt.numpy()), tf.convert_to_tensor(gen.numpy()))
    r = E.result()
    assert list(r) == ["mse", "rmse", "mae", "psnr", "lpips", "ssim"]
    for k in ("mse", "rmse", "mae", "psnr", "ssim"):
        assert abs(r[k] - float(g["same." + k])) <= 1e-6 * abs(r[k])
    assert abs(E.get_progress_bar_info()["img_rgbl1"] - r["mae"]) < 1e-9


def test_reference_test_step_and_predict_step_equal_the_oracle_expectations(tf):
    """migt.py:507-541 as shipped, with the real torch codebook attached: the quantities tests/test_eval_gpu.py holds MIGT.test_step /
    predict_step to (mean loss of call(compute_losses=True), token accuracy beyond the first n_loss_skip views, teacher-forced argmax
    tokens and their decoded images) are what the reference's own steps return."""
    from oracle import make_golden as mg
    vcfg = VQGANConfig(**mg.SMALL_VQ)
    vq = ref_loader.build_reference_vqgan(synth.make_vqgan_state_dict(vcfg, 0), **mg.SMALL_VQ)
    kw = dict(n_layer=2, n_head=4, d_model=64, sequence_size=4, n_embeddings=vcfg.n_embed, token_image_size=8, n_loss_skip=1)
    cfg = MIGTConfig(**kw)
    sd = synth.make_migt_state_dict(cfg, 3)
    model = ref_loader.build_reference_migt(sd, **kw)
    model.codebook_model = ref_loader.ReferenceCodebookNHWC(vq)
    codes = synth.make_codes(3, 4, n_embed=cfg.n_embeddings, side=8, seed=5)
    rel = mo.normalize_cameras(mo.to_relative_cameras(synth.make_cameras(3, 4, seed=6))[0])
    with torch.no_grad():
        o = mo.forward(sd, cfg, dict(input_ids=codes, poses=rel), compute_losses=True)
        for m in model.metrics:
            m.reset_states()
        res = model.test_step((tf.constant(rel.numpy()), codes.clone()))
        ps = model.predict_step((tf.constant(rel.numpy()), codes.clone()))
    assert {"loss", "ce_loss", "acc", "localization_weight", "psnr", "pose_loss", "pose_pos_loss", "pose_ori_loss", "pose_pos_err", "pose_ori_err"} <= set(res)
    assert abs(float(res["loss"]) - float(o["loss"].mean())) < 1e-5 * abs(float(o["loss"].mean()))
    acc = float((o["logits"].argmax(-1)[:, 1:] == codes[:, 1:]).float().mean())
    assert abs(float(res["acc"]) - acc) < 1e-6 and float(res["psnr"]) > 0
    assert torch.equal(torch.as_tensor(ps["latent_code"]).long(), o["logits"].argmax(-1))
    assert tuple(ps["decoded_image"].shape) == (12, 32, 32, 3) and tuple(ps["ground_truth_image"].shape) == (12, 32, 32, 3)

