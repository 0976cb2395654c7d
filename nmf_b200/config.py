"""A small hydra-compatible config loader / instantiator for the render path.

The reference is driven by hydra (`train.py:904-917`, `configs/default.yaml:3-7`): config groups `dataset/`, `model/`,
`field/`, `key=value` overrides, `cfg.model.arch.rf = cfg.field` (train.py:911), and recursive `_target_` /
`_partial_: True` instantiation.  hydra-core / omegaconf are not installed in this image, so this module implements
the subset the path needs with the same semantics:
  * YAML 1.1 reads `1e-3` as a string; OmegaConf reads a float -- scalars are re-parsed accordingly;
  * `_target_` names of the reference (`modules.tensor_nerf.TensorNeRF`, ...) resolve to the mirrors in
    nmf_b200/plugins.py, so reference config files are usable unchanged;
  * `_partial_: True` -> functools.partial; nested configs are instantiated first; lists pass through; NULL -> None.
"""
import copy
import functools
import importlib
import os
import re

import yaml

CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")

TARGETS = {
    "modules.tensor_nerf.TensorNeRF": "nmf_b200.plugins.TensorNeRF",
    "samplers.alphagrid.AlphaGridSampler": "nmf_b200.plugins.AlphaGridSampler",
    "fields.tensoRF.TensorVMSplit": "nmf_b200.plugins.TensorVMSplit",
    "models.microfacet.Microfacet": "nmf_b200.plugins.Microfacet",
    "models.tensorf.TensoRF": "nmf_b200.plugins.PlainTensoRF",
    "brdf_samplers.ggx.GGXSampler": "nmf_b200.plugins.GGXSampler",
    "modules.brdf.MLPBRDF": "nmf_b200.plugins.MLPBRDF",
    "modules.ish.ListISH": "nmf_b200.plugins.ListISH",
    "modules.render_modules.RandHydraMLPDiffuse": "nmf_b200.plugins.RandHydraMLPDiffuse",
    "modules.render_modules.MLPRender_Fea": "nmf_b200.plugins.MLPRender_Fea",
    "modules.integral_equirect.IntegralEquirect": "nmf_b200.plugins.IntegralEquirect",
    "modules.tonemap.SRGBTonemap": "nmf_b200.plugins.SRGBTonemap",
}
_FLOAT = re.compile(r"^[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?$")


class Cfg(dict):
    """dict with attribute access (the part of DictConfig the reference's code uses)."""
    __getattr__ = dict.get
    __setattr__ = dict.__setitem__


def _fix(node):
    if isinstance(node, dict):
        return Cfg({k: _fix(v) for k, v in node.items()})
    if isinstance(node, list):
        return [_fix(v) for v in node]
    if isinstance(node, str):
        if node in ("NULL", "Null", "null", "~"):
            return None
        if _FLOAT.match(node):
            return float(node)
    return node


def load_yaml(path):
    with open(path) as f:
        return _fix(yaml.safe_load(f) or {})


def to_plain(node):
    if isinstance(node, dict):
        return Cfg({k: to_plain(v) for k, v in node.items()})
    if isinstance(node, (list, tuple)):
        return [to_plain(v) for v in node]
    return node


def _set(cfg, dotted, value):
    keys = dotted.split(".")
    for k in keys[:-1]:
        cfg = cfg.setdefault(k, Cfg())
    cfg[keys[-1]] = value


def compose(overrides=(), config_dir=CONFIG_DIR, config_name="default"):
    """hydra.compose: defaults list + `group=choice` / `a.b.c=value` overrides, then the rf <- field splice."""
    cfg = load_yaml(os.path.join(config_dir, f"{config_name}.yaml"))
    groups = {}
    for d in cfg.pop("defaults", []):
        if isinstance(d, dict):
            groups.update(d)
    plain = []
    for o in overrides:
        k, v = o.split("=", 1)
        if k in groups and os.path.isdir(os.path.join(config_dir, k)):
            groups[k] = v
        else:
            plain.append((k, _fix(yaml.safe_load(v))))
    for g, choice in groups.items():
        cfg[g] = load_yaml(os.path.join(config_dir, g, f"{choice}.yaml"))
    for k, v in plain:
        _set(cfg, k, v)
    if "model" in cfg and "field" in cfg:
        cfg["model"]["arch"]["rf"] = copy.deepcopy(cfg["field"])          # train.py:911
    return cfg


def expand_multirun(overrides):
    """hydra `-m` (README.md:10-17 of the reference: `python train.py -m dataset=ficus,helmet,toaster`): every override
    whose value is a top-level comma list sweeps over its items; the jobs are the cartesian product, run sequentially,
    in hydra's order (the last sweep varies fastest).  Bracketed values (`a=[1,2]`) are lists, not sweeps."""
    import itertools
    axes = []
    for o in overrides:
        k, v = o.split("=", 1)
        items, depth, cur = [], 0, ""
        for ch in v:
            if ch in "[{(":
                depth += 1
            elif ch in "]})":
                depth -= 1
            if ch == "," and depth == 0:
                items.append(cur)
                cur = ""
            else:
                cur += ch
        items.append(cur)
        axes.append([f"{k}={it}" for it in items])
    return [list(job) for job in itertools.product(*axes)]


def save(cfg, path):
    """OmegaConf.save(config=cfg, f=path) (train.py:485): the composed run config as YAML, loadable by load_yaml."""
    def plain(node):
        if isinstance(node, dict):
            return {k: plain(v) for k, v in node.items()}
        if isinstance(node, (list, tuple)):
            return [plain(v) for v in node]
        return node
    with open(path, "w") as f:
        yaml.safe_dump(plain(cfg), f, sort_keys=False)


def _resolve(name):
    name = TARGETS.get(name, name)
    mod, _, attr = name.rpartition(".")
    return getattr(importlib.import_module(mod), attr)


def instantiate(node, **extra):
    """hydra.utils.instantiate (recursive)."""
    if isinstance(node, (list, tuple)):
        return [instantiate(v) for v in node]
    if not isinstance(node, dict):
        return node
    if "_target_" not in node:
        return Cfg({k: instantiate(v) for k, v in node.items()})
    kwargs = {k: instantiate(v) for k, v in node.items() if k not in ("_target_", "_partial_")}
    kwargs.update(extra)
    fn = _resolve(node["_target_"])
    return functools.partial(fn, **kwargs) if node.get("_partial_", False) else fn(**kwargs)


def build_model(overrides=(), aabb=None, near_far=None, config_dir=CONFIG_DIR):
    """`hydra.utils.instantiate(cfg.model.arch)(aabb=..., near_far=...)` (train.py:239)."""
    import torch
    cfg = compose(overrides, config_dir)
    if aabb is None:
        s = float(cfg.dataset.get("aabb_scale", 1.0) or 1.0)
        aabb = torch.tensor([[-1.5, -1.5, -1.5], [1.5, 1.5, 1.5]]) * s
    near_far = near_far if near_far is not None else cfg.dataset.near_far
    return instantiate(cfg.model.arch)(aabb=aabb, near_far=near_far), cfg
