"""Generate tests/golden/pipeline_*.npz with the REFERENCE's own inference code (build container only).

  python tests/golden/make_pipeline_golden.py [case name ...]

The reference's model.py (WCTModel graph, build_decoder), ops.py (wct_tf, adain, pad_reflect,
Conv2DReflect), vgg_normalised.py (vgg_from_t7) and torchfile.py are imported UNMODIFIED from
/root/reference and evaluated eagerly over tests/golden/np_tf1.py (a NumPy stand-in for the few
TensorFlow/Keras calls they make -- TensorFlow itself is not installable offline).  Each fixture holds
the inputs, the float64 run of that code ("exact statement" of the reference algorithm), its
float32 run (reference numerics sample) and the per-level tensors of the float64 run.

Weights are NOT stored (too large): they are wct_tf_b200.weights.make_synthetic_weights(seed),
re-derived by the tests; `wsum` guards against generator drift.  The VGG weights reach the
reference code through a Torch7 file written by tests/t7_writer.py and read by the reference's
torchfile.py; decoder weights are served to the reference's Conv2D layers by name, which is what
tf.train.Saver.restore does in wct.py:45-56.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import nets, ref_ops  # noqa: E402
from tests.golden import np_tf1  # noqa: E402
from tests.t7_writer import write_vgg_t7  # noqa: E402
from wct_tf_b200 import weights as W  # noqa: E402

ALL = ["relu5_1", "relu4_1", "relu3_1", "relu2_1", "relu1_1"]
CASES = [
    # name, relu_targets, content HxW, style HxW, alpha, adain, seed, swap5: None | ss_alpha | (ss_alpha, ss_patch_size, ss_stride)
    ("wct5_a06", ALL, (48, 64), (64, 48), 0.6, False, 11, None),
    ("wct_21_41_a10", ["relu2_1", "relu4_1"], (40, 56), (48, 48), 1.0, False, 12, None),     # any order / subset (README.md:46)
    ("wct_31_11_odd_a08", ["relu3_1", "relu1_1"], (37, 45), (41, 50), 0.8, False, 13, None),  # odd sizes: pool 'same' + upsample grow the frame
    ("adain4_a07", ALL[1:], (48, 48), (40, 56), 0.7, True, 14, None),
    ("swap5_51_31_a08", ["relu5_1", "relu3_1"], (96, 112), (112, 96), 0.8, False, 15, 0.6),   # --swap5: style swap at relu5_1 (ops.py:145-278), WCT at relu3_1
    # --ss-patch-size 5 --ss-stride 2: the 9x11 relu5_1 encoding of a 144x176 frame is tiled exactly ((4-1)*2+5 = 11, (3-1)*2+5 = 9),
    # i.e. the size utils.swap_filter_fit would crop to; 4 x 3 style patches
    ("swap5_p5s2_51_21_a07", ["relu5_1", "relu2_1"], (144, 176), (176, 144), 0.7, False, 16, (0.6, 5, 2)),
]


def weight_checksum(w):
    tot = 0.0
    for l in w["vgg"]:
        tot += float(np.abs(l["weight"].astype(np.float64)).sum() + np.abs(l["bias"].astype(np.float64)).sum())
    for t in sorted(w["decoders"]):
        for l in w["decoders"][t]:
            tot += float(np.abs(l["kernel"].astype(np.float64)).sum() + np.abs(l["bias"].astype(np.float64)).sum())
    return tot


def main():
    tmp = tempfile.mkdtemp()
    with np_tf1.reference_modules() as ref:
        only = set(sys.argv[1:])                                   # optional: regenerate just the named cases
        for name, targets, hwc, hws, alpha, adain, seed, ss in CASES:
            if only and name not in only:
                continue
            if ss is None:
                swap = {}
            elif isinstance(ss, tuple):
                swap = dict(swap5=True, ss_alpha=ss[0], ss_patch_size=ss[1], ss_stride=ss[2])
            else:
                swap = dict(swap5=True, ss_alpha=ss)
            w = W.make_synthetic_weights(seed, relu_targets=targets)
            t7 = os.path.join(tmp, name + ".t7")
            write_vgg_t7(t7, w["vgg"])
            dec = {l["name"]: (l["kernel"], l["bias"]) for t in targets for l in w["decoders"][t]}
            rng = np.random.default_rng(seed)
            content = rng.integers(0, 256, hwc + (3,), dtype=np.uint8)
            style = rng.integers(0, 256, hws + (3,), dtype=np.uint8)
            c01, s01 = content[None] / 255.0, style[None] / 255.0          # wct.py:60-64
            out64, lv64 = np_tf1.run_reference(ref, c01, s01, t7, dec, targets, alpha, adain, np.float64, **swap)
            out32, _ = np_tf1.run_reference(ref, c01, s01, t7, dec, targets, alpha, adain, np.float32, **swap)
            # well-posedness of the vector: no covariance eigenvalue near the 1e-5 cut at any level
            _, info = nets.pipeline(content, style, w, targets, alpha=alpha, adain=adain, semantics="tf", dtype=np.float64,
                                    return_info=True, **swap)
            ks = []
            if not adain:
                for inf in info:
                    assert ref_ops.spectral_gap_ok(inf["wc"]) and ref_ops.spectral_gap_ok(inf["ws"]), (name, inf["relu"])
                    ks.append((inf["k_c"], inf["k_s"]))
                    if "margin" in inf:                    # style swap: the arg-max must be decided by a clear margin everywhere
                        assert inf["margin"].min() > 1e-3, (name, float(inf["margin"].min()))
            arrays = dict(content=content, style=style, alpha=np.float64(alpha), adain=np.bool_(adain), seed=np.int64(seed),
                          relu_targets=np.array(targets), out_ref_fp64=out64, out_ref_fp32=out32, wsum=np.float64(weight_checksum(w)),
                          swap5=np.bool_(ss is not None), ss_alpha=np.float64(swap.get("ss_alpha", 0.6)),
                          ss_patch_size=np.int64(swap.get("ss_patch_size", 3)), ss_stride=np.int64(swap.get("ss_stride", 1)),
                          k=np.array(ks, dtype=np.int64).reshape(-1, 2))
            for i, (enc, dec_in, decoded) in enumerate(lv64):
                arrays["lvl%d_decoder_input" % i] = dec_in.astype(np.float32)
                arrays["lvl%d_decoded" % i] = decoded
            np.savez_compressed(os.path.join(HERE, "pipeline_%s.npz" % name), **arrays)
            print(name, "out", out64.shape, "ref fp32-vs-fp64 %.2e" % np.abs(out32 - out64).max(), "k", ks,
                  "range %.3f..%.3f" % (out64.min(), out64.max()))


if __name__ == "__main__":
    main()
