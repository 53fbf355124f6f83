"""Keras HDF5 weight file (the reference's `saved_model_*.h5`, model.py:1047-1060 ModelCheckpoint / load_weights 1157-1196)
-> .npz keyed by the Keras variable names, which MaskYOLO.load_weights of this package reads.

Needs h5py, so it runs where the reference's own environment is available -- NOT in this repository's image, where h5py
is absent; it is therefore untested here.  Layout it relies on (Keras 2.x `save_weights`): root attribute `layer_names`;
one group per layer with attribute `weight_names` (e.g. b'conv1/kernel:0', for the nested model
b'conv_dw_7/depthwise_kernel:0' inside group 'yolo_model'); one dataset per weight name.  A full-model file
(`model.save`) keeps the same structure under the group 'model_weights'.

    python scripts/h5_to_npz.py saved_model.h5 saved_model.npz
"""
import sys

import numpy as np


def convert(src, dst):
    import h5py
    out = {}
    with h5py.File(src, "r") as f:
        g = f["model_weights"] if "model_weights" in f and "layer_names" not in f.attrs else f
        for layer in g.attrs["layer_names"]:
            layer = layer.decode() if isinstance(layer, bytes) else layer
            grp = g[layer]
            for wn in grp.attrs["weight_names"]:
                wn = wn.decode() if isinstance(wn, bytes) else wn
                key = wn.split(":")[0]
                parts = key.split("/")
                key = "/".join(parts[-2:]) if len(parts) > 2 else key      # drop a nested-model prefix
                if key in out:
                    raise ValueError("duplicate variable name %s" % key)
                out[key] = np.asarray(grp[wn], dtype=np.float32)
    with open(dst, "wb") as fh:
        np.savez(fh, **out)
    return out


if __name__ == "__main__":
    if len(sys.argv) != 3:
        sys.exit(__doc__)
    res = convert(sys.argv[1], sys.argv[2])
    print("wrote %s: %d variables, %d values" % (sys.argv[2], len(res), sum(v.size for v in res.values())))
