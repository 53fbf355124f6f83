"""Checkpoint files of the B200 path: a flat {Keras variable name: array} mapping.

The reference saves Keras HDF5 weight files (model.py:1157-1196: `load_weights(filepath, by_name, exclude)`, nested
`yolo_model` group).  h5py is not a dependency of this package; what is read and written here are containers of the SAME
key space -- the variable names of the reference's graph (`conv1/kernel`, `conv_dw_7_bn/moving_mean`,
`myolo_mask_conv1/bias`, ...; SURVEY 10.3, pinned by tests/golden/graph_fixture.json):

    .pt / .pth / anything else   torch.save of the dict (what MaskYOLO.train writes after every epoch)
    .npz                         numpy archive, one array per variable
    .safetensors                 safetensors file, one tensor per variable
    .h5 / .hdf5 / .keras         Keras 2.x HDF5 weight files, read and written by myolo.h5lite (pure Python restatement of
                                 the HDF5 subset h5py produces for save_weights; anything outside it raises with the feature's
                                 name -- scripts/h5_to_npz.py converts such a file where h5py is installed)

Keras appends ':0' to variable names inside HDF5 files and prefixes nested-model variables with the model name;
`normalise_key` strips both so that converted files load unchanged."""
import os

import numpy as np
import torch

_H5 = (".h5", ".hdf5")


def normalise_key(key: str) -> str:
    """'yolo_model/conv_dw_7/depthwise_kernel:0' -> 'conv_dw_7/depthwise_kernel'"""
    key = key.split(":")[0]
    parts = key.split("/")
    return "/".join(parts[-2:]) if len(parts) > 2 else key


def read_checkpoint(path, exclude=None) -> dict:
    """-> {variable name: float32 CPU tensor}; `exclude`: layer names to drop (model.py:1170-1180)."""
    p = str(path)
    ext = os.path.splitext(p)[1].lower()
    if ext in _H5:
        from . import h5lite
        raw = {k: torch.from_numpy(v) for k, v in h5lite.read_keras_weights(p).items()}
    elif ext == ".npz":
        with np.load(p) as z:
            raw = {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}
    elif ext == ".safetensors":
        from safetensors.torch import load_file
        raw = load_file(p, device="cpu")
    else:
        raw = torch.load(p, map_location="cpu", weights_only=True)      # tensors only: no pickled code runs
        if not isinstance(raw, dict):
            raise TypeError("%s does not hold a {variable name: tensor} mapping" % p)
    raw = select(raw, exclude)         # on the names as stored: `exclude` names top-level layers, incl. the nested 'yolo_model'
    return {normalise_key(k): torch.as_tensor(v).to(torch.float32) for k, v in raw.items()}


def write_checkpoint(path, state: dict) -> None:
    p = str(path)
    ext = os.path.splitext(p)[1].lower()
    cpu = {k: torch.as_tensor(v).detach().to("cpu", torch.float32).contiguous() for k, v in state.items()}
    if ext in _H5:
        from . import h5lite
        h5lite.write_keras_weights(p, keras_layers(cpu))
    elif ext == ".npz":
        with open(p, "wb") as f:                       # np.savez would append '.npz' to other names; keep the given path
            np.savez(f, **{k: v.numpy() for k, v in cpu.items()})
    elif ext == ".safetensors":
        from safetensors.torch import save_file
        save_file(cpu, p)
    else:
        torch.save(cpu, p)


def select(state: dict, exclude=None) -> dict:
    """`exclude`: layer names whose variables are dropped (model.py:1170-1180 filters the top-level layers by name; the
    nested 'yolo_model' is one such layer, and naming one of ITS layers works too)."""
    if not exclude:
        return dict(state)
    ex = set(exclude)
    return {k: v for k, v in state.items() if not (set(k.split(":")[0].split("/")[:-1]) & ex)}


_NESTED = "yolo_model"          # model.py:281-292: conv_dw_7..14, conv_pw_7..14 (+ BN) and conv_23 live in this nested Model


def keras_layers(state: dict):
    """[(layer name, [(stored weight name, array), ...])] in the order and grouping Keras' save_weights uses: one entry per
    top-level layer, the second-phase YOLO layers gathered under the nested model."""
    import re
    nested_rx = re.compile(r"(conv_(dw|pw)_(7|8|9|1[0-4])(_bn)?|conv_23)")
    layers, index = [], {}
    for k, v in state.items():
        lname = k.split("/")[0]
        owner = _NESTED if nested_rx.fullmatch(lname) else lname
        if owner not in index:
            index[owner] = len(layers)
            layers.append((owner, []))
        layers[index[owner]][1].append((k + ":0", np.asarray(v)))
    return layers
