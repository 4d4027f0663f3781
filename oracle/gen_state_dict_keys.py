"""Writes tests/golden/reference_state_dict.json: parameter/buffer names and shapes of the REAL reference
NeRSembleNGPModel (imported from /root/reference on the oracle/tp stubs), default recipe hparams.
    python -m oracle.gen_state_dict_keys"""
import json, os, sys
sys.path.insert(0, "/root/reference/src")
import torch
from oracle.tp import install_stubs
install_stubs()
from oracle.gen_golden import build_reference_model
from oracle import pipeline as pl

P = pl.random_params(n_timesteps=4, log2_hashmap_size=12)
m = build_reference_model(P, 4, 12)
sd = {k: list(v.shape) for k, v in m.state_dict().items()}
groups = {k: len(v) for k, v in m.get_param_groups().items()}
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "reference_state_dict.json")
json.dump({"state_dict": sd, "param_groups": groups}, open(out, "w"), indent=1)
print(len(sd), "keys ->", out)
