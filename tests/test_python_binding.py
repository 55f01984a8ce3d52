"""ode_b200/ode.py: the reference's Python surface (bindings/python/ode.pyx) on the classic C API.  One script
(tests/pyode_app.py, tutorial3-style) runs on the unmodified reference library and on the B200 library."""
import os
import numpy as np
import pytest
from parity_util import ROOT
from ode_b200 import ode
import pyode_app


def _ref(prec):
    p = os.path.join(ROOT, "oracle", "_ref", "libode_ref_%s.so" % prec)
    return p if os.path.exists(p) else None


def test_surface_matches_ode_pyx_names():
    """classes / functions / constants a script written for the reference's module expects"""
    for n in ("World", "Body", "Mass", "JointGroup", "Joint", "BallJoint", "HingeJoint", "SliderJoint", "UniversalJoint", "Hinge2Joint",
              "FixedJoint", "ContactJoint", "AMotor", "LMotor", "GeomObject", "SpaceBase", "SimpleSpace", "HashSpace", "Space", "GeomSphere",
              "GeomBox", "GeomPlane", "GeomCapsule", "GeomCCylinder", "GeomCylinder", "GeomRay", "Contact", "collide", "areConnected", "InitODE", "CloseODE", "environment",
              "ParamLoStop", "ParamHiStop2", "ParamFMax3", "ParamSuspensionERP", "paramVel", "ContactBounce", "ContactApprox1", "ContactSoftCFM",
              "AMotorUser", "AMotorEuler", "Infinity"):
        assert hasattr(ode, n), n
    assert (ode.ParamHiStop2, ode.ParamFMax3, ode.ContactApprox1, ode.AMotorEuler) == (257, 517, 0x7000, 1)
    for cls, names in ((ode.World, "setGravity getGravity setERP setCFM quickStep setQuickStepNumIterations setContactMaxCorrectingVel "
                                   "setContactSurfaceLayer setAutoDisableFlag setLinearDamping setAngularDamping impulseToForce"),
                       (ode.Body, "setPosition getPosition setRotation getRotation getQuaternion setQuaternion setLinearVel getLinearVel "
                                  "setAngularVel getAngularVel setMass getMass addForce addTorque getForce getTorque enable disable isEnabled "
                                  "setGravityMode setKinematic isKinematic vectorToWorld getRelPointPos getNumJoints"),
                       (ode.HingeJoint, "attach getBody setFeedback getFeedback setAnchor getAnchor setAxis getAxis setParam"),
                       (ode.AMotor, "setMode setNumAxes setAxis getAxisRel setAngle getAngle setParam"),
                       (ode.GeomBox, "setBody getBody setPosition getAABB setCollideBits setCategoryBits getLengths setOffsetPosition clearOffset"),
                       (ode.HashSpace, "add remove query getNumGeoms getGeom collide setLevels getLevels")):
        for n in names.split():
            assert callable(getattr(cls, n)), (cls.__name__, n)


@pytest.mark.parametrize("prec", ("single", "double"))
def test_host_side_objects_without_gpu(prec):
    """construction, setters / getters and the loud failure of quickStep on the B200 library without a CUDA device"""
    ode.use(precision=prec)
    w = ode.World()
    w.setGravity((0, -9.81, 0))
    assert np.allclose(w.getGravity(), (0, -9.81, 0))
    w.setERP(0.4)
    assert abs(w.getERP() - 0.4) < 1e-6
    s = ode.HashSpace()
    s.setLevels(-2, 6)
    assert s.getLevels() == (-2, 6)
    b = ode.Body(w)
    m = ode.Mass()
    m.setBox(1000, 1.0, 0.2, 0.2)
    b.setMass(m)
    assert abs(b.getMass().mass - 40.0) < 1e-3
    b.setPosition((1, 2, 3))
    b.setRotation([0, -1, 0, 1, 0, 0, 0, 0, 1])
    assert np.allclose(b.getPosition(), (1, 2, 3)) and np.allclose(b.vectorToWorld((1, 0, 0)), (0, 1, 0))
    g = ode.GeomBox(s, (1.0, 0.2, 0.2))
    g.setBody(b)
    assert g.getBody() is b and len(s) == 1 and s.getGeom(0) is g and s.query(g)
    j = ode.AMotor(w)
    j.attach(b, ode.environment)
    j.setMode(ode.AMotorEuler)
    assert j.getNumAxes() == 3 and j.getBody(0) is b
    with pytest.raises(NotImplementedError):
        w.step(0.01)                     # dWorldStep is outside the subset, and says so
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            w.quickStep(0.01)
    else:                                # dGeomGetAABB runs the AABB kernel (the library has no CPU path)
        assert np.allclose(g.getAABB(), (0.9, 1.1, 1.5, 2.5, 2.9, 3.1), atol=1e-6)


@pytest.mark.parametrize("prec", ("single", "double"))
def test_script_runs_on_the_reference(prec):
    """the binding itself is library-agnostic: the tutorial3-style script on the unmodified reference library (CPU)"""
    if not _ref(prec):
        pytest.skip("oracle/_ref not built")
    ode.use(_ref(prec))
    out = pyode_app.run(ode, nsteps=120)
    assert out["ncontacts"] > 100 and out["nbodies"] >= 6 and out["space_len"] == out["nbodies"] + 2
    assert len(out["ranges"]) >= 100 and abs(out["ranges"][0][3] - 0.1) < 1e-3          # the cart's sensor sees the floor 0.1 below its centre
    assert np.isfinite(np.array(out["log"][-1])).all() and np.allclose(out["hinge_axis"], (0, 0, 1), atol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("prec,space_type", (("single", 1), ("double", 1), ("single", 0)))
def test_script_reference_vs_b200(prec, space_type):
    """same script, reference library vs B200 library: same contact count, trajectories within the stated tolerance (cullPoints and the
    hinge angle call atan2: CUDA libm vs glibc)"""
    if not _ref(prec):
        pytest.skip("oracle/_ref not built")
    ode.use(_ref(prec))
    a = pyode_app.run(ode, nsteps=160, space_type=space_type)
    ode.use(precision=prec)
    b = pyode_app.run(ode, nsteps=160, space_type=space_type)
    tol = 0.0 if prec == "single" else 1e-8           # single: atan2 (cullPoints, hinge angle) is the host libm's algorithm, bit-identical
    for s, (x, y) in enumerate(zip(a["log"], b["log"])):
        d = float(np.abs(np.array(x) - np.array(y)).max())
        assert d <= tol, "state differs by %.3g at step %d" % (d, s)
    assert a["ncontacts"] == b["ncontacts"] and a["nbodies"] == b["nbodies"] and a["seed"] == b["seed"]
    assert [r[:3] for r in a["ranges"]] == [r[:3] for r in b["ranges"]]
    assert np.abs(np.array([r[3] for r in a["ranges"]]) - np.array([r[3] for r in b["ranges"]])).max() <= tol
