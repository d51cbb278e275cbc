from .trainer import Trainer, sampler_config_from_json
