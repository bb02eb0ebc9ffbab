// src/main.cu:10 spells this include with a capital T; the file on disk is taylorGreenScenario.cuh.
// This alias lets the reference's main.cu build on a case-sensitive file system; the scenario itself comes from the
// user's / the reference's scenario directory on the include path.
#include "scenarios/taylorGreen/taylorGreenScenario.cuh"
