/* GLU stand-in for the oracle build (see gl.h in this directory). */
#pragma once
static inline void gluPerspective(double, double, double, double) {}
static inline void gluLookAt(double, double, double, double, double, double, double, double, double) {}
