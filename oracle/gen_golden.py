"""Generate ``tests/golden/*.npz`` by executing the REFERENCE's own AMPConv on CPU.

Run in the build container only (``/root/reference`` must exist):

    python -m oracle.gen_golden

For each case of ``oracle/cases.py::GOLDEN_CASES`` it instantiates the class defined in
``/root/reference/src/ampnet/conv/amp_conv.py`` (loaded by path, unmodified, behind the
PyG stand-in), in float64, assigns the seeded parameters, runs forward + backward of
``(out * d_out).sum()`` and stores inputs and results.  It also stores the answers of the
reference's known-answer script ``synthetic_benchmark/testing_message_passing_pyg.py``.
"""
import io
import os
import sys
from contextlib import redirect_stdout

import numpy as np
import torch

from . import cases, reference_loader

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
MAX_WEIGHT_EDGES = 6   # attention coefficients are stored for the first few edges of big cases


def run_reference(mod, spec, x, edge_index, params, d_out):
    conv = mod.AMPConv(embed_dim=spec["d"], num_heads=spec["h"]).double()
    mha = conv.multi_head_attention
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.from_numpy(params["in_proj_weight"]).double())
        mha.in_proj_bias.copy_(torch.from_numpy(params["in_proj_bias"]).double())
        mha.out_proj.weight.copy_(torch.from_numpy(params["out_proj_weight"]).double())
        mha.out_proj.bias.copy_(torch.from_numpy(params["out_proj_bias"]).double())
    xt = torch.from_numpy(x).double().requires_grad_(True)
    ei = torch.from_numpy(edge_index)
    out = conv(xt, ei)
    (out * torch.from_numpy(d_out).double()).sum().backward()
    return {
        "out": out.detach().numpy(),
        "d_x": xt.grad.numpy(),
        "d_in_proj_weight": mha.in_proj_weight.grad.numpy(),
        "d_in_proj_bias": mha.in_proj_bias.grad.numpy(),
        "d_out_proj_weight": mha.out_proj.weight.grad.numpy(),
        "d_out_proj_bias": mha.out_proj.bias.grad.numpy(),
        "attn_output_weights": conv.attn_output_weights.detach().numpy(),
        "attn_output": conv.attn_output.detach().numpy(),
        "state_dict_keys": np.array(sorted(conv.state_dict().keys())),
    }


def main():
    if not reference_loader.available():
        sys.exit("reference tree not found; goldens can only be regenerated in the build container")
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    mod = reference_loader.load_amp_conv_module()
    for name, spec in cases.GOLDEN_CASES.items():
        x, edge_index, params, d_out = cases.make_inputs(
            spec["n"], spec["e"], spec["f"], spec["d"], spec["h"], graph=spec["graph"],
            seed=1234 + len(name), relu_x=(name == "c2_shape"))
        res = run_reference(mod, spec, x, edge_index, params, d_out)
        big = x.size > 50000 or spec["e"] * spec["f"] ** 2 > 300000
        n_w = min(spec["e"], MAX_WEIGHT_EDGES) if big else spec["e"]
        out_t = np.float32 if big else np.float64
        payload = {
            "n": spec["n"], "e": spec["e"], "f": spec["f"], "d": spec["d"], "h": spec["h"],
            "x": x, "edge_index": edge_index.astype(np.int32), "d_out": d_out,
            **{"param_" + k: v for k, v in params.items()},
            "out": res["out"].astype(out_t), "d_x": res["d_x"].astype(out_t),
            "d_in_proj_weight": res["d_in_proj_weight"], "d_in_proj_bias": res["d_in_proj_bias"],
            "d_out_proj_weight": res["d_out_proj_weight"], "d_out_proj_bias": res["d_out_proj_bias"],
            "weight_edges": np.arange(n_w, dtype=np.int32),
            "attn_output_weights": res["attn_output_weights"][:n_w].astype(out_t),
            "attn_output": res["attn_output"][:n_w].astype(out_t),
            "state_dict_keys": res["state_dict_keys"],
        }
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **payload)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1e6:.2f} MB)")

    # the reference's only pinned numbers: mean aggregation on a 5-node star
    kat = reference_loader.load_message_passing_kat()
    answers = {}
    for flag in (False, True):
        buf = io.StringIO()
        with redirect_stdout(buf):
            kat.main(include_self_loop=flag)
        answers[flag] = buf.getvalue()
    with open(os.path.join(GOLDEN_DIR, "message_passing_kat.txt"), "w") as fh:
        fh.write("# stdout of /root/reference/synthetic_benchmark/testing_message_passing_pyg.py main() under the PyG stand-in\n")
        fh.write("# expected by the script's own comments (:37-40): 6,6,6 without / 5.4,5.4,5.4 with the extra edge\n")
        for flag in (False, True):
            fh.write(answers[flag])
    print(answers)


if __name__ == "__main__":
    main()
