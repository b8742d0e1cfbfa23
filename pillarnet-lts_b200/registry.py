"""det3d-style plugin mechanism: class registries + build_from_cfg + a python-module config loader.

Mirrors det3d/utils/registry.py:6-78 (Registry.register_module, build_from_cfg pops `type` and calls
cls(**args)), det3d/models/registry.py:3-11 (the five model registries), det3d/models/builder.py:34-54
(build_reader/backbone/neck/head/detector) and det3d/torchie/utils/config.py:12-29,77-100
(Config.fromfile -> attribute-access dict), so that configs/pillarnet/*.py construct unchanged.
"""
import importlib.util
import os
import sys


class ConfigDict(dict):
    """attribute-access dict (the role addict.Dict plays for det3d's ConfigDict)."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    @staticmethod
    def wrap(obj):
        if isinstance(obj, dict):
            return ConfigDict({k: ConfigDict.wrap(v) for k, v in obj.items()})
        if isinstance(obj, (list, tuple)):
            return type(obj)(ConfigDict.wrap(v) for v in obj)
        return obj


class Config(ConfigDict):
    @staticmethod
    def fromfile(filename):
        filename = os.path.abspath(filename)
        name = "_pn_cfg_" + os.path.basename(filename)[:-3]
        spec = importlib.util.spec_from_file_location(name, filename)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules.pop(name, None)
        d = {k: v for k, v in vars(mod).items()
             if not k.startswith("__") and not isinstance(v, type(sys))}
        cfg = Config(ConfigDict.wrap(d))
        return cfg


class Registry:
    def __init__(self, name):
        self._name = name
        self._module_dict = {}

    @property
    def name(self):
        return self._name

    def get(self, key):
        return self._module_dict.get(key, None)

    def register_module(self, cls):
        if cls.__name__ in self._module_dict:
            raise KeyError(f"{cls.__name__} is already registered in {self._name}")
        self._module_dict[cls.__name__] = cls
        return cls


def build_from_cfg(cfg, registry, default_args=None):
    assert isinstance(cfg, dict) and "type" in cfg
    args = dict(cfg)
    obj_type = args.pop("type")
    if isinstance(obj_type, str):
        obj_cls = registry.get(obj_type)
        if obj_cls is None:
            raise KeyError(f"{obj_type} is not in the {registry.name} registry")
    else:
        obj_cls = obj_type
    if default_args is not None:
        for k, v in default_args.items():
            args.setdefault(k, v)
    return obj_cls(**args)


READERS = Registry("reader")
BACKBONES = Registry("backbone")
NECKS = Registry("neck")
HEADS = Registry("head")
DETECTORS = Registry("detector")
SECOND_STAGE = Registry("second_stage")     # det3d/models/registry.py:9-11 (Pillar R-CNN)
ROI_HEAD = Registry("roi_head")
POINT_HEAD = Registry("point_head")


def build_reader(cfg):
    return build_from_cfg(cfg, READERS)


def build_backbone(cfg):
    return build_from_cfg(cfg, BACKBONES)


def build_neck(cfg):
    return build_from_cfg(cfg, NECKS)


def build_head(cfg):
    return build_from_cfg(cfg, HEADS)


def build_second_stage_module(cfg):
    return build_from_cfg(cfg, SECOND_STAGE)


def build_point_head(cfg):
    return build_from_cfg(cfg, POINT_HEAD)


def build_roi_head(cfg):
    return build_from_cfg(cfg, ROI_HEAD)


def build_detector(cfg, train_cfg=None, test_cfg=None):
    return build_from_cfg(cfg, DETECTORS, dict(train_cfg=train_cfg, test_cfg=test_cfg))
