// joint.h — 6-row weld joint, API of alxspiker/avbd-demo3d source/joint.h:14-47.  Rows are solved on the GPU
// (csrc/avbd_forces.cuh); the virtuals evaluate the same __host__ __device__ row functions on the host copy.
#pragma once
#include "solver.h"

struct Joint : Force {
    vec3 rA, rB;
    quat initialRelativeOrientation;
    float angularStiffness, angularMotor, angularFracture;
    Joint(Solver* solver, Rigid* bodyA, Rigid* bodyB, const vec3& localAnchorA, const vec3& localAnchorB,
          float linearStiffness = FLT_MAX, float angularStiffness = FLT_MAX, float motor = 0.0f, float fracture = FLT_MAX);
    Joint(Solver* solver, Rigid* bodyB, const vec3& worldAnchor, float linearStiffness = FLT_MAX, float angularStiffness = FLT_MAX,
          float motor = 0.0f, float fracture = FLT_MAX);
    int getRowCount() const override { return 6; }
    bool initialize() override { return true; }
    void computeConstraint(float dt) override;
    void computeDerivatives(vec3& J_linear, vec3& J_angular, const Rigid* body, int row) const override;
    void draw() const override {}
    int deviceKind() const override { return 0; }
};
