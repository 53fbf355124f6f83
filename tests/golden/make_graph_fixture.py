"""Extract the structural pins of the reference's shipped TensorBoard GraphDef.

Run HERE (the container that has /root/reference); the output JSON is committed so
tests never read /root/reference at run time.

    python tests/golden/make_graph_fixture.py

Source: /root/reference/example/shapes/tf_graph/events.out.tfevents.1545939845.* (the only
machine-checkable artefact the reference ships; see SURVEY.md header).
"""
import glob
import json
import os
import sys

from tensorboard.backend.event_processing.event_file_loader import RawEventFileLoader
from tensorboard.compat.proto import event_pb2, graph_pb2

SRC = glob.glob("/root/reference/example/shapes/tf_graph/events.out.tfevents.*")[0]
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "graph_fixture.json")


def main():
    graph = None
    scalars = {}
    for raw in RawEventFileLoader(SRC).Load():
        ev = event_pb2.Event.FromString(raw)
        if ev.graph_def:
            graph = graph_pb2.GraphDef.FromString(ev.graph_def)
        if ev.HasField("summary"):
            for v in ev.summary.value:
                if v.HasField("simple_value"):
                    scalars.setdefault(v.tag, []).append([ev.step, v.simple_value])
    assert graph is not None
    variables = {}
    attrs = {}
    placeholders = {}
    for n in graph.node:
        if n.op in ("VariableV2", "Variable"):
            shp = [d.size for d in n.attr["shape"].shape.dim]
            variables[n.name] = shp
        elif n.op == "Placeholder":
            if n.name.startswith("Placeholder"):
                continue  # Keras weight-loading feeds, not model inputs
            shp = [d.size for d in n.attr["shape"].shape.dim]
            placeholders[n.name] = shp
        elif n.op in ("Conv2D", "DepthwiseConv2dNative", "Conv2DBackpropInput") and "gradients" not in n.name:
            attrs[n.name] = {
                "op": n.op,
                "strides": list(n.attr["strides"].list.i),
                "padding": n.attr["padding"].s.decode(),
                "data_format": n.attr["data_format"].s.decode(),
            }
        elif n.op == "FusedBatchNorm" and "gradients" not in n.name:
            attrs[n.name] = {"op": n.op, "epsilon": n.attr["epsilon"].f,
                             "is_training": bool(n.attr["is_training"].b)}
        elif n.op == "CropAndResize" and "gradients" not in n.name:
            attrs[n.name] = {"op": n.op, "method": n.attr["method"].s.decode(),
                             "extrapolation_value": n.attr["extrapolation_value"].f}
    # constants that pin paddings / crop sizes / thresholds
    consts = {}
    import numpy as np
    from tensorboard.util import tensor_util
    for n in graph.node:
        if n.op == "Const" and "gradients" not in n.name and (
                n.name.endswith("Pad/paddings") or n.name.endswith("crop_size")
                or "GreaterEqual/y" in n.name or "Less/y" in n.name
                or n.name.startswith("Adam/") ):
            try:
                val = tensor_util.make_ndarray(n.attr["value"].tensor)
                if val.size <= 16:
                    consts[n.name] = np.asarray(val).tolist()
            except Exception:
                pass
    # keep only model variables (not optimizer slots)
    model_vars = {k: v for k, v in variables.items() if not k.startswith("training/")}
    out = {
        "source": os.path.basename(SRC),
        "num_nodes": len(graph.node),
        "variables": model_vars,
        "num_optimizer_slot_vars": len(variables) - len(model_vars),
        "placeholders": placeholders,
        "op_attrs": attrs,
        "consts": consts,
        "scalars": scalars,
    }
    with open(OUT, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print("wrote", OUT, "vars", len(model_vars), "attrs", len(attrs), "consts", len(consts))


if __name__ == "__main__":
    sys.exit(main())
