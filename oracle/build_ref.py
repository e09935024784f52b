"""Recipe for oracle/_ref: the UNMODIFIED reference package, installed from /root/reference where it lies.

    python oracle/build_ref.py

TEST / BENCH INFRASTRUCTURE ONLY.  The reference (domkirke/acids_transforms 0.1.3) is pure Python on top of torch /
torchaudio; "building" it is an offline `pip install --no-deps --target oracle/_ref` of a scratch copy of the read-only
source tree (pip writes egg-info next to setup.py).  Nothing of it enters the repository's history: oracle/_ref/ is
git-ignored, but it is NOT gpurun-ignored, so it travels to the GPU box where /root/reference does not exist.
Consumers: `bench.py --impl reference`, bench.py's `cpu_baseline` and `sub.eager_cuda` legs (all through
bench.import_reference, which adds the one shim the reference needs in this image: a stub `turtle` module for
acids_transforms/transforms/misc.py:1).  When oracle/_ref is absent those legs fall back to oracle/torch_port.py and
say `kind: "port"`.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("ACIDS_REFERENCE", "/root/reference")
DEST = os.path.join(HERE, "_ref")


def build(force=False):
    if not os.path.isdir(REF_SRC):
        return None                                   # GPU box: use what travelled with the snapshot
    if os.path.isdir(os.path.join(DEST, "acids_transforms")) and not force:
        return DEST
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REF_SRC, src, ignore=shutil.ignore_patterns(".git", "__pycache__"))
        shutil.rmtree(DEST, ignore_errors=True)
        subprocess.run([sys.executable, "-m", "pip", "install", "--quiet", "--no-index", "--no-build-isolation", "--no-deps",
                        "--target", DEST, src], check=True)
    return DEST


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
