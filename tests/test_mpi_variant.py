"""Row f3 of SURVEY.md section 8: the MPI-INF-3DHP variant (capf_b200.mpi.VolumetricTriangulationNet) against fixtures
generated from the unmodified reference tree ContextPose_mpi/ (oracle/gen_golden_mpi.py) and against the oracle."""
import contextlib
import io
import json
import os

import numpy as np
import pytest
import torch

import capf_b200
import capf_oracle
import interp
import protocol
from capf_b200 import mpi, program
from conftest import GOLDEN, load_golden, rel_l2

CASES = ["mpi_hrnet32_b2_128x96", "mpi_hrnet48_b2_128x96"]


def _model(backbone, wseed, precision="fp32"):
    cfg = mpi.make_mpi_config(backbone)
    with contextlib.redirect_stdout(io.StringIO()):
        m = mpi.VolumetricTriangulationNet(cfg, precision=precision).eval()
    w = protocol.make_weights([(k, tuple(v.shape)) for k, v in m.state_dict().items()], wseed)
    m.load_state_dict(w, strict=True)
    return m, w, cfg


def _meta(name):
    g = load_golden(name)
    return g, json.loads(str(g["meta"]))


@pytest.mark.parametrize("backbone", ["hrnet_32", "hrnet_48"])
def test_mpi_state_dict_matches_reference_manifest(backbone):
    """Same keys and shapes as ContextPose_mpi's VolumetricTriangulationNet.state_dict() (no context_blocks;
    embed_dim_ratio 64 / 96): run_3dhp.py's bare-state_dict checkpoints load with strict=True."""
    with open(os.path.join(GOLDEN, "state_dict_manifest_mpi.json")) as f:
        want = {k: tuple(s) for k, s in json.load(f)[backbone]}
    m, _, _ = _model(backbone, 0)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    assert not any("context_blocks" in k for k in got)


@pytest.mark.parametrize("name", CASES)
def test_mpi_oracle_and_program_match_reference_fixture(name):
    """CPU: the oracle's restatement (capf_oracle.mpi_forward) and the op program executed by the reference interpreter
    (same memory plan and packed weights the GPU uses) both reproduce the reference's output (b,3,1,17,1) and its
    in-place crop normalisation."""
    g, meta = _meta(name)
    B, H, W = meta["B"], meta["H"], meta["W"]
    m, w, cfg = _model(meta["backbone"], meta["wseed"])
    images, kp2d, crop = protocol.make_inputs(B, H, W, meta["iseed"])
    c = crop.clone()
    out, second = capf_oracle.mpi_forward(w, cfg.model.backbone, int(cfg.model.poseformer.depth), images, kp2d, c)
    assert second is None and tuple(out.shape) == (B, 3, 1, 17, 1)
    assert rel_l2(out, g["out"]) < 1e-5 and np.array_equal(c.numpy(), g["crop_after"])
    prog = program.build_forward_program(meta["backbone"], m.backbone.cfg, m._pf_cfg, {k: tuple(v.shape) for k, v in w.items()},
                                         B, H, W, "fp32", variant="mpi")
    assert not any("context_blocks" in op.tag for op in prog.ops)
    it = interp.Interp(prog, w)
    it.t(prog.inputs["images"]).copy_(images)
    it.t(prog.inputs["kp2d"]).copy_(kp2d.reshape(-1, 2))
    it.t(prog.inputs["ref"]).copy_(torch.from_numpy(g["crop_after"]).reshape(-1, 2))
    it.run()
    y = it.t(prog.outputs["out"]).view(B, 1, 17, 3)
    assert rel_l2(y.reshape(B, 1, 17, 3, 1).permute(0, 3, 1, 2, 4), g["out"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("fp16", 2.5e-3)])
def test_mpi_forward_on_gpu_matches_reference_fixture(name, precision, tol):
    """B200: VolumetricTriangulationNet.forward through the C ABI == the reference's output (x, None), crop mutated
    identically; fp16 mode within the documented 16-bit tolerance."""
    g, meta = _meta(name)
    B, H, W = meta["B"], meta["H"], meta["W"]
    m, _, _ = _model(meta["backbone"], meta["wseed"], precision)
    m = m.cuda()
    images, kp2d, crop = protocol.make_inputs(B, H, W, meta["iseed"])
    c = crop.clone().cuda()
    with torch.no_grad():
        out, second = m(images.cuda(), kp2d.cuda(), c)
    assert second is None and tuple(out.shape) == (B, 3, 1, 17, 1) and out.is_contiguous()
    assert np.array_equal(c.cpu().numpy(), g["crop_after"])
    assert rel_l2(out.cpu(), g["out"]) < tol


@pytest.mark.gpu
def test_mpi_batch_of_one_works_where_the_reference_fails():
    """The reference squeezes the batch dimension away at B = 1 (pose_dformer.py:240 `.squeeze()`) and raises; frames
    are independent here, so B = 1 equals the first frame of a larger batch."""
    m, _, _ = _model("hrnet_32", 0)
    m = m.cuda()
    images, kp2d, crop = protocol.make_inputs(2, 128, 96, 21)
    with torch.no_grad():
        full, _ = m(images.cuda(), kp2d.cuda(), crop.clone().cuda())
        one, _ = m(images[:1].cuda(), kp2d[:1].cuda(), crop[:1].clone().cuda())
    assert rel_l2(one.cpu(), full[:1].cpu()) < 1e-5
