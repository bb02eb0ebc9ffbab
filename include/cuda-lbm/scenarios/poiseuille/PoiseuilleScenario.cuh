// case alias for src/main.cu:13 (see scenarios/taylorGreen/TaylorGreenScenario.cuh)
#include "scenarios/poiseuille/poiseuilleScenario.cuh"
