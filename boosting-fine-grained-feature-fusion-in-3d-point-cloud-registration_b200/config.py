"""KPConv option sets of the reference's three shipped configurations.

Values restated from ``conf/3dmatch.yaml:26-54``, ``conf/mcd.yaml`` (identical kpconv_options) and
``conf/modelnet.yaml:35-58`` of the reference.  The reference flattens its YAML sections into one
``EasyDict`` (``utils/misc.py:10-29``) that is handed to ``Preprocessor`` / ``KPFEncoder``; the
object returned here offers the same attribute + ``.get()`` access.
"""
from __future__ import annotations

import copy


class AttrDict(dict):
    """dict with attribute access (stand-in for easydict.EasyDict, which is not a dependency)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as exc:
            raise AttributeError(name) from exc

    def __setattr__(self, name, value):
        self[name] = value


_COMMON = dict(
    aggregation_mode="sum",
    fixed_kernel_points="center",
    in_feats_dim=1,
    in_points_dim=3,
    deform_radius=5.0,
    KP_extent=2.0,
    KP_influence="linear",
    use_batch_norm=True,
    batch_norm_momentum=0.02,
    modulated=False,
    num_kernel_points=15,
)

_RESNET_4 = ["simple", "resnetb", "resnetb_strided", "resnetb", "resnetb", "resnetb_strided",
             "resnetb", "resnetb", "resnetb_strided", "resnetb", "resnetb"]

_CONFIGS = {
    "3dmatch": dict(_COMMON, num_layers=4, neighborhood_limits=[40, 40, 40, 40],
                    first_subsampling_dl=0.025, first_feats_dim=128, conv_radius=2.5,
                    overlap_radius=0.0375, architecture=_RESNET_4, d_embed=512),
    "mcd": dict(_COMMON, num_layers=4, neighborhood_limits=[40, 40, 40, 40],
                first_subsampling_dl=0.025, first_feats_dim=128, conv_radius=2.5,
                overlap_radius=0.0375, architecture=_RESNET_4, d_embed=512),
    "modelnet": dict(_COMMON, num_layers=2, neighborhood_limits=[50, 50],
                     first_subsampling_dl=0.03, first_feats_dim=512, conv_radius=2.75,
                     overlap_radius=0.04,
                     architecture=["simple", "resnetb", "resnetb", "resnetb_strided", "resnetb",
                                   "resnetb"], d_embed=256),
}


def kpconv_config(name: str, **overrides) -> AttrDict:
    """Return the kpconv option set ``name`` ('3dmatch' | 'mcd' | 'modelnet') as an AttrDict."""
    cfg = AttrDict(copy.deepcopy(_CONFIGS[name]))
    cfg.update(overrides)
    return cfg
