"""In-tree build of libdronesim_b200.so with nvcc for sm_100a (B200).

    python -m mujoco_drone_b200.build [--force] [--verbose]

The .so lands next to this file (git-ignored, but it travels with the gpurun snapshot).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdronesim_b200.so")
SOURCES = ["dsim_kernels.cu", "dsim_policy_mlp.cu", "dsim_policy_fp32.cu"]
HEADERS = ["dsim_device.cuh", "dsim_obs_reward.cuh", "dsim_params.cuh", "dsim_step.cuh", "dsim_step_x2.cuh", "dsim_packed.cuh", "dsim_contact.cuh", "dsim_policy.cuh", os.path.join("..", "..", "include", "dronesim_b200.h")]
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared"]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libdronesim_b200.so")


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return all(os.path.getmtime(d) <= t for d in deps)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libdronesim_b200.so")
    return OUT


def build_variant(name, defines, verbose=False):
    """kernel-variant experiments (tools/sweep_variants.sh): mujoco_drone_b200/variants/<name>.so compiled with extra -D flags;
    loaded through DSIM_LIB, never by default"""
    vdir = os.path.join(HERE, "variants")
    os.makedirs(vdir, exist_ok=True)
    out = os.path.join(vdir, name + ".so")
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-D" + d for d in defines] + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building variant " + name)
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:                       # python -m mujoco_drone_b200.build --variant name DEF1 DEF2=3 ...
        k = sys.argv.index("--variant")
        print(build_variant(sys.argv[k + 1], [a for a in sys.argv[k + 2:] if not a.startswith("--")], verbose="--verbose" in sys.argv))
        sys.exit(0)
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
