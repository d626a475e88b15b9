"""Process-global configuration object mirroring the keys the reference model reads from its
module-level `args` EasyDict (src/utils/args.py): `num_classes` (CamRaDepth.py:38),
`supervised_seg` / `unsupervised_seg` (:42-43), `input_channels` (:45), `groupnorm_divisor`
(simplified_attention.py:22; utils.py:209).  `set_model(name)` applies args.py:156-168.
"""
from __future__ import annotations


class Args(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


MODELS = ["base", "base (rgb)", "supervised_seg", "unsupervised_seg", "sup_unsup_seg", "sup_unsup_seg (rgb)"]

args = Args(num_classes=21, supervised_seg=False, unsupervised_seg=False, input_channels=7,
            groupnorm_divisor=16, model="base", learning_rate=6e-05, max_depth=100, update_interval=1,
            hashtags_prefix="####################################")


def set_model(name: str) -> Args:
    assert name in MODELS, "Model type invalid"
    args.model = name
    args.supervised_seg = name in ["sup_unsup_seg", "sup_unsup_seg (rgb)", "supervised_seg"]
    args.unsupervised_seg = name in ["sup_unsup_seg", "sup_unsup_seg (rgb)", "unsupervised_seg"]
    args.input_channels = 3 if name in ["base (rgb)", "sup_unsup_seg (rgb)"] else 7
    return args
