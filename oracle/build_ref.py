"""Stages the reference's OWN Python sources for the render path under oracle/_ref/ (git-ignored, travels to the GPU box
with gpurun like a built .so) so that GPU-side checks can run the unmodified reference code:

  * tests/test_dropin_reference_gpu.py builds the reference's real GaussianModel / GaussianLearner / FeaturePlanes and
    renders it through splatco_b200 (drop-in proof, INTEGRATION.md options A and B);
  * bench.py's `gpu_baseline` times the reference's own generate_neural_gaussians on the same B200.

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing is copied into tracked files and nothing under splatco_b200/ imports it.
Run here (the authoring container, where /root/reference exists):  python oracle/build_ref.py
"""
import os
import shutil
import sys

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = [
    "gaussian_renderer/__init__.py", "gaussian_renderer/network_gui.py",
    "scene/__init__.py", "scene/gaussian_model.py", "scene/grids.py", "scene/embedding.py", "scene/cameras.py",
    "scene/dataset_readers.py", "scene/colmap_loader.py",
    "utils/general_utils.py", "utils/graphics_utils.py", "utils/system_utils.py", "utils/sh_utils.py",
    "utils/loss_utils.py", "utils/grid_utils.py", "utils/entropy_models.py", "utils/camera_utils.py", "utils/image_utils.py",
    "arguments/__init__.py", "train.py", "render.py",
]


def main():
    if not os.path.isdir(REF):
        print("oracle/build_ref.py: /root/reference is absent; nothing staged")
        return 0
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    n = 0
    for f in FILES:
        src = os.path.join(REF, f)
        if not os.path.exists(src):
            continue
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    with open(os.path.join(DST, "STAGED_FROM"), "w") as fh:
        fh.write(f"{REF}: {n} files staged by oracle/build_ref.py (unmodified copies; not tracked)\n")
    print(f"oracle/build_ref.py: staged {n} reference files under {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
