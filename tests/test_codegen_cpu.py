"""CPU suite (authoring container only: needs /root/reference and the built parser-free driver,
host/codegen/gala_b200_codegen): the retargeted generator emits bindings instead of kernel text, prepares the
graphs on the GPU, and its fusion pass (decided on the IR, applied to whole statements) fuses exactly the GAT
layers whose four nodes share one graph slot and the transform + attention-projection groups."""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CG = os.path.join(ROOT, "gala-gnn-acceleration-language_b200", "host", "codegen")
DRIVER = os.path.join(CG, "gala_b200_codegen")

pytestmark = pytest.mark.skipif(not (os.path.exists(DRIVER) and os.path.isdir("/root/reference/src/codegen")),
                                reason="needs /root/reference and host/codegen/gala_b200_codegen (make -C host/codegen)")


def emit(tmp_path, model, mode, tile, *flags):
    out = tmp_path / f"{model}_{mode}_{'_'.join(f.strip('-') for f in flags)}"
    out.mkdir()
    r = subprocess.run([DRIVER, model, "Reddit", "602", "41", str(tile), mode, str(out) + "/", ROOT, *flags],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    return (out / "gala.cu").read_text(), r.stdout


def test_retargeted_program_carries_bindings_not_kernel_text(tmp_path):
    src, _ = emit(tmp_path, "gcn", "inference", 37000)
    assert '#include "gala_b200_torch.h"' in src
    assert "GALA_B200_DEFINE_AGGREGATE_TILED(" in src
    assert "__global__" not in src                      # no kernel text is pasted into the program
    ref, _ = emit(tmp_path, "gcn", "inference", 37000, "--reference")
    assert "__global__" in ref and "gala_b200" not in ref


def test_gat_layers_are_fused_only_when_all_nodes_share_a_slot(tmp_path):
    src, log = emit(tmp_path, "gat", "inference", 370000)
    # layer 1 absorbs the torch::relu that follows it, layer 2 has none
    assert len(re.findall(r"res = gala_b200::gat_layer_AutoGrad::apply\(res, attenL, attenR, 0, 0\.2, true\);", src)) == 1
    assert len(re.findall(r"res = gala_b200::gat_layer_AutoGrad::apply\(res, attenL_2, attenR_2, 0, 0\.2, false\);", src)) == 1
    assert "non_lnr_op_softmax_AutoGrad::apply" not in src and "fused 2 GAT layer(s)" in log
    assert "torch::relu" not in src
    # dense side: every Linear retargeted, layer-1 transform + projections one call, layer-2 projections folded
    assert "torch::nn::Linear" not in src and src.count("gala_b200::Linear") == 12
    assert "std::tie(res, attenL, attenR) = gala_b200::linear_att(fc0, efc0, efc1, t_iden);" in src
    assert "std::tie(attenL_2, attenR_2) = gala_b200::folded_att(fc1, efc2, efc3, res);" in src
    assert "res = fc1->forward(res);" in src                      # the classifier after the aggregation stays a call
    # graphs are prepared on the device: none of the reference's CPU headers is included
    assert '#include "gala_b200_host.h"' in src and "../src/formats" not in src and "../tests/common.h" not in src
    host, _ = emit(tmp_path, "gat", "inference", 370000, "--host-formats")
    assert "../src/formats/csrc_matrix.h" in host and "gala_b200_host.h" not in host
    cublas, _ = emit(tmp_path, "gat", "inference", 370000, "--torch-linear")
    assert "gala_b200::Linear" not in cublas and cublas.count("torch::nn::Linear") == 12
    plain, _ = emit(tmp_path, "gat", "inference", 370000, "--no-fuse")
    assert "gat_layer_AutoGrad::apply" not in plain and plain.count("non_lnr_op_softmax_AutoGrad::apply") == 2
    assert "gala_b200::Linear" not in plain and "linear_att" not in plain       # one-to-one bindings only
    # the train driver aggregates over sub-graph slots 1 / 2 while the logits use slot 0: left as emitted
    train, _ = emit(tmp_path, "gat", "train", 370000)
    assert "gat_layer_AutoGrad::apply" not in train and train.count("non_lnr_op_softmax_AutoGrad::apply") == 2


def test_fusion_pass_fails_loudly_on_unexpected_statements(tmp_path):
    """The decision is taken on the IR; if the statements the reference emitted for those nodes do not parse as
    expected, code generation stops (exit 3) instead of silently keeping the slow sequence.  Simulated by asking the
    driver to corrupt one emitted statement before the pass runs."""
    out = tmp_path / "bad"
    out.mkdir()
    r = subprocess.run([DRIVER, "gat", "Reddit", "602", "41", "370000", "inference", str(out) + "/", ROOT,
                        "--corrupt-forward-for-test"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 3 and "unexpected shape" in r.stderr


def test_other_models_keep_their_statements_and_retarget_the_linears(tmp_path):
    for model, mode, n_linear in (("gcn", "inference", 4), ("gin", "train", 4), ("sage", "train", 8)):
        src, log = emit(tmp_path, model, mode, 37000)
        assert f"retargeted {n_linear} torch::nn::Linear" in log and "torch::nn::Linear" not in src
        assert "gat_layer_AutoGrad" not in src and "linear_att" not in src


def test_sampling_and_sage_programs_generate(tmp_path):
    src, _ = emit(tmp_path, "gcn", "inference", 37000, "--sample", "20")
    assert re.search(r"GALA_B200_DEFINE_AGGREGATE_TILED\(\w+, false, 20\)", src)
    sage, _ = emit(tmp_path, "sage", "train", 37000)
    assert "GALA_B200_DEFINE_AGGREGATE_TILED(aggregate_node_mul_sum_direct_coarse2_call, false, 0)" in sage
