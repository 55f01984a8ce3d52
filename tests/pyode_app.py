"""A headless script in the style of the reference's bindings/python/demos/tutorial3.py (boxes dropped on a plane, y up, a near-callback
that calls ode.collide() and creates ContactJoints, contactgroup.empty() after the step) plus a hinged pendulum with a stop and an
LMotor-driven box; written against whatever module object `ode` it is handed (the reference-style API of ode_b200/ode.py).  As in
tests/classic_app.py the callback buffers the pairs and the contacts are created in geom-creation order, so the constraint order does
not depend on a space's internal traversal order; world.quickStep replaces tutorial3's world.step (QuickStep is the path)."""
from math import cos, sin, pi


def create_box(ode, world, space, density, lx, ly, lz):
    body = ode.Body(world)
    M = ode.Mass()
    M.setBox(density, lx, ly, lz)
    body.setMass(M)
    body.shape, body.boxsize = "box", (lx, ly, lz)
    geom = ode.GeomBox(space, lengths=body.boxsize)
    geom.setBody(body)
    return body, geom


def run(ode, nsteps=240, space_type=1):
    ode.randSetSeed(4711)           # the one call that is not in ode.pyx: QuickStep's row shuffles draw from the global dRand
    world = ode.World()
    world.setGravity((0, -9.81, 0))
    world.setERP(0.8)
    world.setCFM(1e-5)
    space = ode.Space(space_type)
    floor = ode.GeomPlane(space, (0, 1, 0), 0)
    order = {floor._id(): 0}
    bodies, geoms = [], [floor]
    contactgroup = ode.JointGroup()

    def add(body, geom):
        order[geom._id()] = len(order)
        bodies.append(body)
        geoms.append(geom)

    # a pendulum on a hinge with stops, hanging from the environment
    arm, g = create_box(ode, world, space, 500, 0.1, 0.6, 0.1)
    arm.setPosition((1.5, 1.2, 0.0))
    add(arm, g)
    hinge = ode.HingeJoint(world)
    hinge.attach(arm, ode.environment)
    hinge.setAnchor((1.5, 1.5, 0.0))
    hinge.setAxis((0, 0, 1))
    hinge.setParam(ode.ParamLoStop, -0.5)
    hinge.setParam(ode.ParamHiStop, 0.5)
    arm.setAngularVel((0, 0, 3.0))
    # a box pushed along x by a linear motor
    cart, g = create_box(ode, world, space, 800, 0.3, 0.2, 0.3)
    cart.setPosition((-1.5, 0.1, 0.5))
    add(cart, g)
    lm = ode.LMotor(world)
    lm.attach(cart, ode.environment)
    lm.setNumAxes(1)
    lm.setAxis(0, 0, (1, 0, 0))
    lm.setParam(ode.ParamVel, 0.6)
    lm.setParam(ode.ParamFMax, 30.0)

    # a rolling wheel (cylinder on the plane) and a downward-looking range sensor (ray) riding on the cart
    wheel = ode.Body(world)
    M = ode.Mass()
    M.setCylinder(600, 3, 0.25, 0.12)
    wheel.setMass(M)
    wheel.setPosition((0.0, 0.25, -1.5))
    wheel.setAngularVel((0, 0, -4.0))
    wg = ode.GeomCylinder(space, 0.25, 0.12)
    wg.setBody(wheel)
    wg.setCategoryBits(2)
    wg.setCollideBits(1)                  # meets the floor only (floor category 1): keeps boxes away, the cylinder-box collider is not built
    floor.setCategoryBits(1)
    add(wheel, wg)
    ray = ode.GeomRay(space, 2.0)
    ray.setBody(cart)
    ray.setOffsetRotation([1, 0, 0, 0, 0, -1, 0, 1, 0])          # local z -> world -y: looks down
    ray.setCategoryBits(4)
    ray.setCollideBits(1 | 8)
    order[ray._id()] = len(order)
    ranges = []

    def drop_object(k):
        body, geom = create_box(ode, world, space, 1000, 1.0, 0.2, 0.2)
        geom.setCategoryBits(8)
        geom.setCollideBits(1 | 4 | 8)
        body.setPosition((0.05 * sin(1.7 * k), 3.0, 0.05 * cos(2.3 * k)))
        theta = 2 * pi * ((0.618 * k) % 1.0)
        ct, st = cos(theta), sin(theta)
        body.setRotation([ct, 0., -st, 0., 1., 0., st, 0., ct])
        add(body, geom)

    pairs = []

    def near_callback(args, geom1, geom2):
        pairs.append(tuple(sorted((geom1, geom2), key=lambda g: order[g._id()])))

    ncontacts, log = 0, []
    dt = 1.0 / 50 / 4
    for step in range(nsteps):
        if step % 24 == 0 and len(bodies) < 10:
            drop_object(len(bodies))
        del pairs[:]
        space.collide((world, contactgroup), near_callback)
        for geom1, geom2 in sorted(pairs, key=lambda p: (order[p[0]._id()], order[p[1]._id()])):
            if ode.areConnected(geom1.getBody(), geom2.getBody()):
                continue
            if geom1 is ray or geom2 is ray:                     # sensor: note the range, make no joint
                for c in ode.collide(geom1, geom2):
                    ranges.append((step, order[geom1._id()], order[geom2._id()], c.getContactGeomParams()[2]))
                continue
            for c in ode.collide(geom1, geom2):
                c.setBounce(0.2)
                c.setMu(5000)
                j = ode.ContactJoint(world, contactgroup, c)
                j.attach(geom1.getBody(), geom2.getBody())
                ncontacts += 1
        world.quickStep(dt)
        contactgroup.empty()
        log.append([b.getPosition() + b.getQuaternion() + b.getLinearVel() + b.getAngularVel() for b in bodies])
    return dict(log=log, seed=ode.randGetSeed(), ranges=ranges, ncontacts=ncontacts, nbodies=len(bodies), npairs=len(pairs), space_len=len(space),
                gravity=world.getGravity(), hinge_axis=hinge.getAxis(), mass=bodies[0].getMass().mass)
