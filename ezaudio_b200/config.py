"""Config wire format: the reference's YAML / JSON files are read as-is (ckpts/ezaudio-xl.yml, ckpts/ezaudio-l.yml,
ckpts/controlnet/energy_l.yml, ckpts/vae/config.json; loader src/utils/utils.py:7-17 incl. its `!include` tag).  When no
file is given the shipped hyper-parameters are used from ezaudio_b200.synth (same values)."""
from __future__ import annotations

import copy
import json
import os

import yaml

from . import synth

DIFF = dict(num_train_timesteps=1000, beta_schedule="scaled_linear", beta_start=0.00085, beta_end=0.012,
            prediction_type="v_prediction", rescale_betas_zero_snr=True, timestep_spacing="trailing", clip_sample=False)
AUTOENCODER = dict(name="stable_vae", dim=128, sr=24000, latent_sr=50, q_first=True, scale=1.0, shift=0.0)

BUILTIN = {
    "s3_xl": dict(model_name="EzAudio-XL", model=synth.XL_MODEL, autoencoder=AUTOENCODER,
                  text_encoder=dict(model="google/flan-t5-xl", max_length=100, cfg=0.1), diff=DIFF),
    "s3_l": dict(model_name="EzAudio-L", model=synth.L_MODEL, autoencoder=AUTOENCODER,
                 text_encoder=dict(model="google/flan-t5-large", max_length=100, cfg=0.1), diff=DIFF),
}
BUILTIN_CONTROLNET = {
    "energy": dict(model_name="EzAudio-L", model=synth.L_MODEL, controlnet=synth.CONTROLNET,
                   conditioner=dict(condition_type="energy", hop_size=240, window_size=1920, padding="reflect", min_db=-60, norm=True),
                   autoencoder=AUTOENCODER, text_encoder=dict(model="google/flan-t5-large", max_length=100, cfg=0.1), diff=DIFF),
}


def load_yaml_with_includes(yaml_file: str):
    def include(loader, node):
        with open(os.path.join(os.path.dirname(yaml_file), loader.construct_scalar(node))) as f:
            return yaml.load(f, Loader=yaml.FullLoader)

    yaml.add_constructor("!include", include, Loader=yaml.FullLoader)
    with open(yaml_file) as f:
        return yaml.load(f, Loader=yaml.FullLoader)


def load_params(name: str, config_path=None, table=None):
    if config_path is not None:
        return load_yaml_with_includes(config_path)
    table = BUILTIN if table is None else table
    if name not in table:
        raise KeyError(f"unknown model_name {name!r}; known: {sorted(table)}")
    return copy.deepcopy(table[name])


def load_vae_encoder_config(path=None):
    if path is None:
        return copy.deepcopy(synth.VAE_ENCODER)
    with open(path) as f:
        return json.load(f)["model"]["encoder"]["config"]


def load_vae_decoder_config(path=None):
    if path is None:
        return copy.deepcopy(synth.VAE_DECODER)
    with open(path) as f:
        return json.load(f)["model"]["decoder"]["config"]
