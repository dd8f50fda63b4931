"""CPU: the product's parameter inventory (reference state_dict key names + shapes) equals the oracle's
state_dict, so a checkpoint that loads into one loads into the other."""
import pytest
import torch

from hipie_b200.modeling import params
from hipie_oracle import hparams
from hipie_oracle.model import HipieOracle


@pytest.mark.parametrize("name", ["vit_tiny", "r50"])
def test_spec_matches_oracle_state_dict(name):
    hp = hparams.get(name)
    if name == "r50":
        hp.update(enc_layers=1, dec_layers=1, md_enc_layers=1, md_dec_layers=1, num_queries=20, md_queries=10,
                  bert=dict(vocab=1000, hidden=768, layers=1, heads=12, inter=3072, max_pos=512))
    oracle = HipieOracle(hp)
    sd = oracle.state_dict()
    spec = params.build_spec(hp)
    canon = {}
    for k, v in sd.items():
        if k.endswith("position_ids") or k.endswith("token_type_ids"):
            continue
        canon.setdefault(params.canonical_name(k, hp), tuple(v.shape))
    missing = sorted(set(canon) - set(spec.shapes))
    extra = sorted(set(spec.shapes) - set(canon))
    assert not missing, f"in oracle/reference but not in spec: {missing[:10]}"
    assert not extra, f"in spec but not in oracle: {extra[:10]}"
    for k, shp in canon.items():
        assert spec.shapes[k] == shp, (k, spec.shapes[k], shp)


def test_random_state_dict_loads_into_oracle():
    hp = hparams.get("vit_tiny")
    sd = params.random_state_dict(hp, seed=1)
    oracle = HipieOracle(hp)
    own = oracle.state_dict()
    full = {}
    for k in own:
        c = params.canonical_name(k, hp)
        full[k] = sd[c] if c in sd else own[k]
    oracle.load_state_dict(full, strict=True)
