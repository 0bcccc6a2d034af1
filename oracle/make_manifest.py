"""Container-only: instantiate the unmodified reference engine and record its state_dict layout
(key -> shape) as udifftext_b200/manifests/{full,tiny}.json.  TEST INFRASTRUCTURE / one-off generator."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_import  # noqa: E402
from udifftext_b200 import synth  # noqa: E402


def small_overrides(name: str):
    if name == "full":
        return None
    a = synth.ARCH[name]
    unet = {k: a["unet"][k] for k in ("model_channels", "attention_resolutions", "num_res_blocks", "channel_mult", "t_context_dim")}
    label = {k: a["label"][k] for k in ("emb_dim", "n_trans_layers")}
    vae = {k: a["vae"][k] for k in ("ch", "ch_mult", "num_res_blocks")}
    return dict(unet=unet, label=label, vae=vae)


def main():
    for name in sys.argv[1:] or ["tiny", "full"]:
        m = ref_import.build_reference_engine(1234, small_overrides(name))
        man = {k: list(v.shape) for k, v in m.state_dict().items()}
        path = os.path.join(synth.MANIFEST_DIR, f"{name}.json")
        with open(path, "w") as f:
            json.dump(man, f, indent=0)
        print(name, len(man), "tensors", sum(int(__import__('numpy').prod(s)) for s in man.values()) / 1e6, "M params ->", path)


if __name__ == "__main__":
    main()
