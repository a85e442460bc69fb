"""viewformer_b200.metrics.Evaluator on the GPU against the numbers of the REFERENCE's own Evaluator (evaluate/evaluate_transformer.py:22-67
with the metric classes of utils/metrics.py), produced by running those files over oracle/tf_shim.py (oracle/make_golden.py ->
tests/golden/evaluator_reference_shim.npz; reproduced in the container by tests/test_reference_on_shim.py).

What the fixture encodes beyond the formulas: `mse` / `mae` on the 0..255 scale (Keras casts the uint8 images), `ssim` with K1 = 1
(SSIMMetric passes 1 as ssim()'s third positional argument, metrics.py:183), ground truth resized by the dataset rule and the generated
image bilinearly (evaluate_transformer.py:42-45).  Inputs are rebuilt bit for bit from integer ops (oracle/synth.py::make_metric_pair).

(The file name sorts after every other test module on purpose: it was added after the round's last GPU session.)
"""
import os

import numpy as np
import pytest

from oracle import synth
from oracle.make_golden import EVALUATOR_CASES, evaluator_cameras

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", EVALUATOR_CASES, ids=[c[0] for c in EVALUATOR_CASES])
def test_evaluator_reports_the_reference_evaluators_numbers(golden_dir, case):
    from viewformer_b200.metrics import Evaluator
    tag, n, gs, ns, image_size, seed = case
    g = np.load(os.path.join(golden_dir, "evaluator_reference_shim.npz"))
    gt, gen = synth.make_metric_pair(n, gs, ns, seed)
    assert [int(gt.sum()), int(gen.sum())] == g[f"{tag}.input_sums"].tolist()
    ev = Evaluator(image_size)
    ev.update_with_image(gt[:2], gen[:2])               # two updates: the running means accumulate per image
    ev.update_with_image(gt[2:], gen[2:])
    r = ev.result()
    # no resize: integer-exact inputs, only fp32-vs-fp64 rounding of the reference's float32 metrics is left.  With a resize the CUDA
    # kernel may differ from torch's interpolate by 1 LSB in < 0.1 % of the pixels (tests/test_eval_gpu.py), worth < 1e-4 relative here.
    rel = 1e-5 if gs == ns and image_size is None else 2e-3
    for k in ("mse", "rmse", "mae", "psnr", "ssim"):
        want = float(g[f"{tag}.{k}"])
        print(f"[evaluator {tag}] {k}: got {r[k]:.7f} want {want:.7f}")
        assert abs(r[k] - want) <= rel * abs(want) + (2e-5 if k == "ssim" else 0.0), (tag, k, r[k], want)
    if tag == "same":
        # the K1 quirk is visible: the function-level default (image_metrics' ssim) differs from what the Evaluator reports
        from viewformer_b200.metrics import image_metrics
        d = float(image_metrics(gt, gen)["ssim"].mean())
        assert abs(d - float(g["same.ssim_default_k1"])) < 2e-5 and abs(d - r["ssim"]) > 1e-4


def test_evaluator_cameras_and_images_together(golden_dir):
    from viewformer_b200.metrics import Evaluator
    g = np.load(os.path.join(golden_dir, "evaluator_reference_shim.npz"))
    gt, gen = synth.make_metric_pair(5, 64, 64, 21)
    cg, cn = evaluator_cameras()
    ev = Evaluator()
    ev.update_state(ground_truth_cameras=cg, generated_cameras=cn, ground_truth_images=gt, generated_images=gen)
    r = ev.result()
    for k in ("loc-angle", "loc-dist", "loc-angle-med", "loc-dist-med"):
        assert abs(r[k] - float(g["cam." + k])) < 2e-6 * max(1.0, abs(r[k])), k
    assert abs(r["psnr"] - float(g["same.psnr"])) < 1e-5 * r["psnr"]
    info = ev.get_progress_bar_info()
    assert set(info) == {"img_psnr", "cam_loc", "cam_ang"} and abs(info["img_psnr"] - r["psnr"]) < 1e-12


def test_codebook_round_trip_and_its_evaluator(golden_dir):
    """evaluate/evaluate_codebook.py:66-76 (BASELINE configs[0]: encode -> decode round trip) through viewformer_b200.evaluate
    .generate_codebook_predictions, against the oracle chain that the reference script itself reproduces on the CPU
    (tests/test_reference_on_shim.py::test_codebook_evaluation_script_equals_oracle_and_fixture); CodebookEvaluator = the image half."""
    import torch
    from oracle import vqgan_oracle as vo, migt_oracle as mo
    from oracle.make_golden import SMALL_VQ
    from viewformer_b200 import VQGAN, generate_codebook_predictions
    from viewformer_b200.config import VQGANConfig
    from viewformer_b200.metrics import CodebookEvaluator
    vcfg = VQGANConfig(**SMALL_VQ)
    vsd = synth.make_vqgan_state_dict(vcfg, 0)
    cb = VQGAN(vcfg, precision="fp32").load_state_dict(vsd)
    images = synth.make_images_uint8(1, 3, size=vcfg.image_size, seed=51)[0]
    with torch.no_grad():
        codes = vo.encode(vsd, vcfg, mo.images_to_float(images).permute(0, 3, 1, 2).contiguous())[2]
        want = mo.float_to_images(vo.decode_code(vsd, vcfg, codes).permute(0, 2, 3, 1))
    r = generate_codebook_predictions(cb, images)
    assert torch.equal(r["codes"].cpu(), codes) and torch.equal(torch.as_tensor(r["ground_truth_images"]), images)
    assert r["generated_images"].dtype == torch.uint8 and int((r["generated_images"].cpu().int() - want.int()).abs().max()) <= 1
    g = np.load(os.path.join(golden_dir, "evaluator_reference_shim.npz"))
    gt, gen = synth.make_metric_pair(5, 64, 64, 21)
    ev = CodebookEvaluator()
    ev.update_state(gt, gen)
    res = ev.result()
    assert set(res) == {"mse", "rmse", "mae", "psnr", "ssim"}
    for k in res:
        assert abs(res[k] - float(g["same." + k])) <= 1e-5 * abs(res[k]) + (2e-5 if k == "ssim" else 0.0), k
    assert abs(ev.get_progress_bar_info()["img_rgbl1"] - res["mae"]) < 1e-12
