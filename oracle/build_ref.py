"""Recipe for `oracle/_ref/`: the REFERENCE's own connected-components kernel, compiled from where it lies.

TEST INFRASTRUCTURE ONLY.  Source: /root/reference/sam2_train/csrc/connected_components.cu (the pybind module the
reference ships as `sam2_train/_C.so`; that prebuilt file is sm_89 + pre-cxx11 ABI and does not load here).  It is built
with `torch.utils.cpp_extension.load` for sm_100a straight from the read-only reference tree — no source is copied into
this repo — into `oracle/_ref/ref_cc/ref_cc.so` (git-ignored, travels to the GPU box with the snapshot).  nvcc
cross-compiles without a GPU; about 3-4 minutes.  `tests/test_gpu_ref_cc.py` loads the built module on the B200 and
checks `ms2_cc_label` against it bit for bit; the GPU box never reads /root/reference.
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref", "ref_cc")
SRC = "/root/reference/sam2_train/csrc/connected_components.cu"
SO = os.path.join(OUT, "ref_cc.so")


def build_ref_cc(force=False):
    """-> path of the built module, or None when the reference tree is absent (GPU box) and nothing was prebuilt."""
    if os.path.exists(SO) and not force:
        return SO
    if not os.path.exists(SRC):
        return None
    os.makedirs(OUT, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    os.environ.setdefault("MAX_JOBS", "4")
    from torch.utils.cpp_extension import load
    load(name="ref_cc", sources=[SRC], build_directory=OUT, verbose=False, is_python_module=False,
         extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
    return SO if os.path.exists(SO) else None


def load_ref_cc():
    """import the prebuilt module (needs torch imported first); None if it was never built."""
    if not os.path.exists(SO):
        return None
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location("ref_cc", SO)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build_ref_cc(force="--force" in sys.argv))
