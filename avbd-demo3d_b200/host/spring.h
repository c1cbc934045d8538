// spring.h — 1-row distance constraint, API of alxspiker/avbd-demo3d source/spring.h:13-37.
#pragma once
#include "solver.h"

struct Spring : Force {
    vec3 rA, rB;
    float restLength, springStiffness;
    mat3 H_ll;                        // kept for layout parity; unused by the solver upstream too (spring.cpp:87-89)
    Spring(Solver* solver, Rigid* bodyA, Rigid* bodyB, const vec3& localAnchorA, const vec3& localAnchorB, float stiffness, float rest = -1.0f);
    int getRowCount() const override { return 1; }
    bool initialize() override { return true; }
    void computeConstraint(float dt) override;
    void computeDerivatives(vec3& J_linear, vec3& J_angular, const Rigid* body, int row) const override;
    void draw() const override {}
    int deviceKind() const override { return 1; }
};
