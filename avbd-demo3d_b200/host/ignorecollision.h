// ignorecollision.h — 0-row marker, API of alxspiker/avbd-demo3d source/ignorecollision.h:14-22.  On the device it
// becomes an entry of the sorted pair-exclusion list the narrowphase cull consults.
#pragma once
#include "solver.h"

struct IgnoreCollision : Force {
    IgnoreCollision(Solver* solver, Rigid* bodyA, Rigid* bodyB) : Force(solver, bodyA, bodyB) {}
    int getRowCount() const override { return 0; }
    bool initialize() override { return true; }
    void computeConstraint(float) override {}
    void computeDerivatives(vec3&, vec3&, const Rigid*, int) const override {}
    int deviceKind() const override { return 2; }
};
