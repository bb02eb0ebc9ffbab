"""CPU: the header shim (include/cuda-lbm/) compiles — against this repo's scenario files and, where the reference tree is
present (the build container), against the reference's OWN scenario files and its own main.cu, unmodified.

nvcc cross-compiles for sm_100a without a GPU; nothing is executed here (tests/test_shim_gpu.py runs the binaries)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/src"
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
BASE = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O1", "-w", f"-I{ROOT}/include/cuda-lbm"]
LINK = [f"-L{ROOT}/cuda_lbm_b200", "-llbm_b200"]

pytestmark = pytest.mark.skipif(not os.path.exists(NVCC), reason="nvcc not available")


def _lib():
    from cuda_lbm_b200 import _capi
    if not os.path.exists(_capi.LIB_PATH):
        import __graft_entry__ as g
        g.build()


def _compile(args, src, out):
    r = subprocess.run(BASE + args + [src] + LINK + ["-o", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert os.path.getsize(out) > 0


def test_own_scenarios_compile_and_link(tmp_path):
    _lib()
    _compile([f"-I{ROOT}/examples", '-DSCENARIO_HEADER="scenarios/b200_cylinder.cuh"', "-DSCENARIO_TYPE=B200CylinderScenario", "-DNX=256", "-DNY=128"],
             f"{ROOT}/examples/main.cu", str(tmp_path / "cyl"))


def test_scenario_trait_members(tmp_path):
    """Every member of the reference's ScenarioTrait (scenario.cuh:22-78) exists with the reference's defaults."""
    src = tmp_path / "t.cu"
    src.write_text(r'''
#include "scenarios/scenario.cuh"
#include "functors/includes.cuh"
struct B { __host__ __device__ int operator()(int, int) const { return BC_flag::FLUID; } };
using T = ScenarioTrait<DefaultInit<2>, B>;
static_assert(std::is_same<T::InitType, DefaultInit<2>>::value && std::is_same<T::BoundaryType, B>::value, "functor types");
static_assert(std::is_same<T::ValidationType, void>::value && !T::has_analytical_solution, "validation");
static_assert(std::is_same<T::CollisionOp, BGK<2>>::value && std::is_same<T::AdapterOp, NoAdapter>::value, "defaults");
static_assert(T::viscosity == 1.0f / 6.0f && T::tau == 1.0f && T::omega == 1.0f && T::u_max == 0.1f, "constants");
static_assert(T::S[0] == 0.0f && T::S[1] == 1.0f && T::S[3] == 0.0f && T::S[5] == 0.0f && T::S[8] == 1.0f, "S");
static_assert(quadratures == 9 && dimensions == 2, "lattice");
static_assert(BC_flag::BOUNCE_BACK == 1 && BC_flag::ZOU_HE_LEFT == 3 && BC_flag::CYLINDER == 6 && BC_flag::ZG_OUTFLOW == 7 &&
              BC_flag::PRESSURE_OUTLET == 8 && BC_flag::REGULARIZED_INLET_TOP == 9 && BC_flag::REGULARIZED_BOUNCE_BACK == 11 &&
              BC_flag::REGULARIZED_BOUNCE_BACK_CORNER == 12, "BC_flag values are the ABI of lbm_set_flags");
static_assert(CM<2, OptimalAdapter>::lbm_b200_op == LBM_CM_OPTIMAL && CM<2, NoAdapter>::lbm_b200_op == LBM_CM && MRT<2>::lbm_b200_op == LBM_MRT, "ops");
static_assert(viscosity_to_tau(0.1f) == 0.8f, "helpers");
int main() {
    T::update_ts(5.0f); T::add_bodies();
    IBMBody b = create_cylinder(10.f, 10.f, 4.f);
    bool ok = get_vec_index(7, 1) == 15 && get_node_from_coords(3, 2) == 2 * NX + 3 && T::t == 5.0f && T::IBM_bodies.empty() && std::string(T::name()) == "BaseScenario" && b.num_points == 16 && b.points[0] == 14.0f;
    h_ibm_free(b);
    return ok ? 0 : 1;
}
''')
    out = str(tmp_path / "t")
    r = subprocess.run(BASE + ["-DNX=32", "-DNY=16", str(src), "-o", out], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    assert subprocess.run([out]).returncode == 0       # host-only program: no device needed


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("defs", [["-DUSE_TAYLOR_GREEN", "-DPERIODIC_X", "-DPERIODIC_Y"], ["-DUSE_POISEUILLE", "-DPERIODIC_X"], ["-DUSE_LID_DRIVEN"]],
                         ids=["taylorGreen", "poiseuille", "lidDrivenCavity"])
def test_reference_scenario_files_compile_unchanged(tmp_path, defs):
    """src/scenarios/{taylorGreen,poiseuille,lidDrivenCavity} of the reference, as they lie, through examples/main.cu."""
    _lib()
    _compile([f"-I{REF_SRC}"] + defs, f"{ROOT}/examples/main.cu", str(tmp_path / "a.out"))


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
def test_reference_main_cu_compiles_unchanged(tmp_path):
    """The reference's own driver, src/main.cu.  Compiled from a scratch copy: a quoted #include looks next to the including
    file first, so inside src/ it would find the reference's core/lbm.cuh instead of the shim's."""
    _lib()
    shutil.copy(f"{REF_SRC}/main.cu", tmp_path / "main.cu")
    _compile([f"-I{REF_SRC}", "-DUSE_TAYLOR_GREEN", "-DPERIODIC_X", "-DPERIODIC_Y"], str(tmp_path / "main.cu"), str(tmp_path / "a.out"))


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="reference tree not present (GPU box)")
def test_reference_cylinder_scenario_needs_only_its_one_token_fix(tmp_path):
    """flowPastCylinderScenario.cuh names `BGK` without <2>: it does not compile in the reference
    either.  With that one token repaired in a scratch copy the file compiles against the shim; create_cylinder, which the
    reference forgets to declare, comes from scenarios/scenario.cuh."""
    _lib()
    d = tmp_path / "scenarios" / "flowPastCylinder"
    d.mkdir(parents=True)
    for fn in ("flowPastCylinderScenario.cuh", "flowPastCylinderFunctors.cuh"):
        txt = open(f"{REF_SRC}/scenarios/flowPastCylinder/{fn}").read()
        if fn.endswith("Scenario.cuh"):
            assert "    BGK\n" in txt
            txt = txt.replace("    BGK\n", "    BGK<2>\n", 1)
        (d / fn).write_text(txt)
    _compile([f"-I{tmp_path}", "-DUSE_FLOW_PAST_CYLINDER"], f"{ROOT}/examples/main.cu", str(tmp_path / "a.out"))
