"""Model configs — host-side mirror of the reference's ``config.json`` schema.

Field names, defaults and the ``model`` discriminator follow
``viewformer/models/config.py:39-119`` (``ModelConfig``, ``MIGTConfig``, ``VQGANConfig``) and
``viewformer/models/__init__.py:62-78`` (``load_config``) so a reference ``config.json`` parses here
unchanged.  ``localization_weight`` is a reference ``Schedule`` string (``utils/schedules.py``);
only "is it identically zero?" matters on the inference path, so it is kept as a string.
"""
import json
from dataclasses import dataclass, field, fields, asdict, is_dataclass
from typing import List


class ModelNotFoundError(RuntimeError):
    """Same role as viewformer/models/__init__.py:11-12."""


@dataclass
class ModelConfig:
    def __post_init__(self):
        name = type(self).__name__
        assert name.endswith("Config")
        self.model = name[: -len("Config")].lower()

    def asdict(self):
        d = asdict(self)
        d["model"] = self.model
        return d


@dataclass
class MIGTConfig(ModelConfig):
    n_embeddings: int = 1024
    n_head: int = 12
    d_model: int = 768
    dropout: float = 0.1
    n_layer: int = 12
    weight_decay: float = 0.01
    label_smoothing: float = 0.0
    learning_rate: float = 6.4e-4
    batch_size: int = 64
    gradient_clip_val: float = 0.0
    sequence_size: int = 20
    token_image_size: int = 8
    total_steps: int = 300000
    n_loss_skip: int = 4
    augment_poses: str = "relative"          # 'no' | 'relative' | 'simple' | 'advanced'
    use_dynamic_pose_loss: bool = False
    localization_weight: str = "1"           # reference Schedule, kept as its string form
    image_generation_weight: float = 1.0
    pose_multiplier: float = 1.0
    random_pose_multiplier: float = 1.0

    @property
    def model_type(self):
        return "transformer"

    @property
    def use_localization(self):
        """migt.py:268-269: ``not localization_weight.is_zero()``."""
        from .schedules import parse
        return not parse(self.localization_weight).is_zero()


@dataclass
class VQGANConfig(ModelConfig):
    learning_rate: float = 1.584e-3
    embed_dim: int = 256
    n_embed: int = 1024
    z_channels: int = 256
    resolution: int = 256
    in_channels: int = 3
    out_ch: int = 3
    ch: int = 128
    num_res_blocks: int = 2
    ch_mult: List[int] = field(default_factory=lambda: [1, 1, 2, 2, 4])
    attn_resolutions: List[int] = field(default_factory=lambda: [16])
    gradient_clip_val: float = 0.0
    batch_size: int = 352
    image_size: int = 128
    total_steps: int = 200000
    codebook_weight: float = 1.0
    pixelloss_weight: float = 1.0
    perceptual_weight: float = 1.0

    @property
    def stride(self):
        return 2 ** (len(self.ch_mult) - 1)

    @property
    def model_type(self):
        return "codebook"


_CONFIGS = {"migt": MIGTConfig, "vqgan": VQGANConfig}


def supported_config_dict():
    return dict(_CONFIGS)


def load_config(config):
    """dict / JSON path / config object -> config object (models/__init__.py:62-78)."""
    if isinstance(config, ModelConfig):
        return config
    if is_dataclass(config) and not isinstance(config, type) and hasattr(config, "model"):
        # a config object of the reference itself (viewformer/models/config.py dataclasses, handed over by its AutoModel*)
        config = {f.name: getattr(config, f.name) for f in fields(config)}
        config = {k: (str(v) if hasattr(v, "from_str") else v) for k, v in config.items()}
    if isinstance(config, str):
        with open(config) as f:
            config = json.load(f)
    config = dict(config)
    name = config.pop("model", None)
    if name not in _CONFIGS:
        raise ModelNotFoundError(f"Model {name} is not supported")
    cls = _CONFIGS[name]
    known = {f.name for f in fields(cls)}
    kwargs = {}
    for k, v in config.items():
        if k not in known:
            continue                     # the reference ignores nothing, but older configs carry extra keys
        if k == "localization_weight":
            v = str(v)
        kwargs[k] = v
    return cls(**kwargs)
