"""Constructor-argument schema of the hot path = the dataclasses of ref:reformer_tts/model/config.py (which cannot be
imported on Python >= 3.11: mutable dataclass defaults at :43-44,:51-53,:80-84) plus a loader that applies a
reference YAML's ``model:`` section over those defaults, giving the kwargs ``ReformerTTS(**kwargs)`` takes
(ref:reformer_tts/training/wrappers.py:37 splats ``asdict(config.model)`` the same way)."""
from __future__ import annotations

import copy
from typing import Any, Dict

MODEL_DEFAULTS: Dict[str, Any] = {
    "num_mel_coeffs": None, "dict_size": None, "embedding_dim": 512, "pad_base": 128, "scp_encoding_dropout": 0.05,
    "enc_prenet_kwargs": {"dropout": 0.5},
    "enc_reformer_kwargs": {
        "depth": 6, "ff_chunks": 100,
        "attn_kwargs": {"implementation": "reformer_pytorch", "heads": 8, "bucket_size": 64, "n_hashes": 8,
                        "add_local_attn_hash": False, "attn_chunks": 1, "random_rotations_per_head": False,
                        "attend_across_buckets": True, "allow_duplicate_attention": True, "num_mem_kv": 0,
                        "one_value_head": False, "use_full_attn": False, "full_attn_thres": None, "return_attn": False,
                        "post_attn_dropout": 0., "dropout": 0.},
        "ff_kwargs": {"hidden": 2048, "dropout": 0.},
    },
    "dec_prenet_kwargs": {"hidden_size": 256, "dropout": 0.5},
    "dec_reformer_kwargs": {
        "depth": 6, "ff_chunks": 100,
        "attn_kwargs": {"num_heads": 8, "dropout": 0., "bias": True, "add_bias_kv": False, "add_zero_attn": False,
                        "kdim": None, "vdim": None},
        "self_attn_kwargs": None,    # filled below: same schema as enc attn_kwargs
        "ff_kwargs": {"hidden": 2048, "dropout": 0.},
    },
    "postnet_kwargs": {"depth": 4, "dropout": 0.},
}
MODEL_DEFAULTS["dec_reformer_kwargs"]["self_attn_kwargs"] = copy.deepcopy(MODEL_DEFAULTS["enc_reformer_kwargs"]["attn_kwargs"])


def _overlay(base: Dict[str, Any], over: Dict[str, Any], path: str = "model") -> Dict[str, Any]:
    for key, value in over.items():
        if key not in base:
            raise KeyError(f"unknown config key {path}.{key}")      # dacite strict=True in the reference
        if isinstance(base[key], dict) and isinstance(value, dict):
            _overlay(base[key], value, f"{path}.{key}")
        else:
            base[key] = value
    return base


def model_kwargs(overrides: Dict[str, Any]) -> Dict[str, Any]:
    """defaults <- ``overrides`` (the ``model:`` mapping of a reference YAML)."""
    cfg = _overlay(copy.deepcopy(MODEL_DEFAULTS), overrides)
    missing = [k for k in ("num_mel_coeffs", "dict_size") if cfg[k] is None]
    if missing:
        raise KeyError(f"config is missing required model keys {missing}")
    return cfg


def model_kwargs_from_yaml(path: str) -> Dict[str, Any]:
    import yaml

    class _Loader(yaml.SafeLoader):
        pass
    _Loader.add_constructor("!path", lambda loader, node: loader.construct_scalar(node))
    with open(path) as fh:
        doc = yaml.load(fh, Loader=_Loader)
    return model_kwargs(doc["model"])


# The ``model:`` sections of the four reference YAMLs named in BASELINE.json, verbatim values
# (ref:config/baseline.yml:35-52, bucket-size-64-18-06.yml:36-55, huggingface-lsh.yml:35-55, depth-3-15-06.yml:35-57).
_COMMON = {"dict_size": 76, "num_mel_coeffs": 80, "scp_encoding_dropout": 0.05, "pad_base": 256,
           "enc_prenet_kwargs": {"dropout": 0.05}, "dec_prenet_kwargs": {"dropout": 0.05}}
REFERENCE_CONFIGS: Dict[str, Dict[str, Any]] = {
    "baseline": dict(_COMMON, enc_reformer_kwargs={"depth": 3},
                     dec_reformer_kwargs={"depth": 3, "self_attn_kwargs": {"bucket_size": 128}},
                     postnet_kwargs={"depth": 2, "dropout": 0.1}),
    "bucket-size-64-18-06": dict(_COMMON, enc_reformer_kwargs={"attn_kwargs": {"post_attn_dropout": 0.15}},
                                 dec_reformer_kwargs={"self_attn_kwargs": {"post_attn_dropout": 0.15}, "attn_kwargs": {"dropout": 0.15}},
                                 postnet_kwargs={"depth": 2, "dropout": 0.3}),
    "huggingface-lsh": dict(_COMMON, enc_reformer_kwargs={"depth": 3, "attn_kwargs": {"implementation": "huggingface_transformers"}},
                            dec_reformer_kwargs={"depth": 3, "self_attn_kwargs": {"implementation": "huggingface_transformers", "bucket_size": 128}},
                            postnet_kwargs={"depth": 2, "dropout": 0.1}),
    "depth-3-15-06": dict(_COMMON, enc_reformer_kwargs={"depth": 3, "attn_kwargs": {"post_attn_dropout": 0.1}},
                          dec_reformer_kwargs={"depth": 3, "self_attn_kwargs": {"bucket_size": 128, "post_attn_dropout": 0.1},
                                               "attn_kwargs": {"dropout": 0.1}},
                          postnet_kwargs={"depth": 2, "dropout": 0.2}),
}
# per-GPU batch sizes BASELINE.json / the YAMLs quote for those configs
REFERENCE_BATCH = {"baseline": 4, "bucket-size-64-18-06": 20, "huggingface-lsh": 12, "depth-3-15-06": 64}


def reference_model_kwargs(name: str) -> Dict[str, Any]:
    return model_kwargs(copy.deepcopy(REFERENCE_CONFIGS[name]))
