"""Generates tests/golden/rank_ctr_layout.json by RUNNING the reference's own slot-slicing code
(/root/reference/rank/ctr/base_model.py: SingleSlot + BaseModel.__init__, pure-Python integer bookkeeping)
on its shipped model_parameter.json, with tensorflow / tensornet replaced by inert stubs (neither is
installable offline; the slicing logic does not depend on them).

Golden INPUT  : the feature table of the shipped config in compact form
                [name, slot_ids, emb_size, bias_type | null] for sparse / sequence / dense features.
Golden OUTPUT : max_embed_size, the sorted sparse slot list, every emb_structure_input slice
                (slot, start, end) in order, the bias slices per bias_type, the gate slices.

    python tools/gen_rank_ctr_golden.py        (only in the container that has /root/reference)
"""
import importlib.util
import json
import os
import sys
import types

REF = "/root/reference/rank/ctr"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "rank_ctr_layout.json")


class Sliced:
    """Stands in for an embedding tensor: records the column slice taken from it."""

    def __init__(self, slot, sl=None):
        self.slot, self.sl = slot, sl
        self.shape = (None, 0)

    def __getitem__(self, key):
        rows, cols = key
        return Sliced(self.slot, (cols.start, cols.stop))


class Feature:
    def __init__(self, feature_id=None, feature_slot=None, sparse=True, **kw):
        self.feature_id = feature_id

    def __lt__(self, other):
        return self.feature_id < other.feature_id


def stub_modules():
    tf = types.ModuleType("tensorflow")
    tf.feature_column = types.SimpleNamespace(embedding_column=lambda col, dimension, combiner: ("emb", col, dimension))
    tn = types.ModuleType("tensornet")
    tn.feature_column = types.SimpleNamespace(FeatureSlot=lambda s: ("slot", s), Feature=Feature,
                                              category_column=lambda key, bucket_size: key)
    captured = {}

    class EmbeddingFeatures:
        def __init__(self, cols, opt, name=None):
            self.cols = cols
            captured["dimension"] = cols[0][2]

        def __call__(self, inputs):
            return {k: Sliced(k) for k in inputs}

    def Input(name=None, feature=None, shape=None, dtype=None, sparse=None, **kw):
        return Sliced(str(getattr(feature, "feature_id", name)))

    tn.layers = types.SimpleNamespace(Input=Input, EmbeddingFeatures=EmbeddingFeatures)
    tn.core = types.SimpleNamespace(Adam=lambda **kw: ("adam", kw))
    sys.modules["tensorflow"], sys.modules["tensornet"] = tf, tn
    return captured


def main():
    captured = stub_modules()
    spec = importlib.util.spec_from_file_location("ref_base_model", os.path.join(REF, "base_model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    cfg = json.load(open(os.path.join(REF, "model_parameter.json")))
    m = mod.BaseModel(cfg)
    fs = cfg["feature_slot"]
    compact = {
        "sparse_feature": [[k, v["slot_id"], v["emb_size"], v.get("bias_type") if "bias" in v else None,
                            ("bias" in v) and ("bias_type" not in v)] for k, v in fs["sparse_feature"].items()],
        "sequence_feature": [[k, v["slot_id"], v["emb_size"]] for k, v in fs["sequence_feature"].items()],
        "dense_feature": [[k, v["slot_id"]] for k, v in fs["dense_feature"].items()],
    }
    out = {
        "input": compact,
        "max_embed_size": captured["dimension"],
        "sparse_slots": sorted(k for k in m.inputs if k not in m.dense_inputs),
        "dense_slots": sorted(m.dense_inputs),
        "structure": [[s.slot, s.sl[0], s.sl[1]] for s in m.emb_structure_input],
        "bias": {t: [[s.slot, s.sl[0], s.sl[1]] for s in lst] for t, lst in m.emb_bias_input.items()},
        "gate": [[s.slot, s.sl[0], s.sl[1]] for s in m.emb_gate_input],
    }
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    json.dump(out, open(OUT, "w"), separators=(",", ":"))
    print(OUT, "structure", len(out["structure"]), "bias", {k: len(v) for k, v in out["bias"].items()},
          "gate", len(out["gate"]), "max_embed_size", out["max_embed_size"], "slots", len(out["sparse_slots"]))


if __name__ == "__main__":
    main()
