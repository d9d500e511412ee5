"""Installs the unmodified reference into ``baseline/_ref`` (offline, no dependencies pulled):

    python baseline/install_ref.py [/root/reference]

The source tree is read-only and setuptools writes an egg-info next to setup.py, so the install runs from a copy in /tmp.
"""
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))


def install(src: str = "/root/reference") -> str:
    dst = os.path.join(HERE, "_ref")
    if not os.path.isdir(src):
        raise FileNotFoundError(src)
    tmp = tempfile.mkdtemp(prefix="torchctr_ref_")
    try:
        copy = os.path.join(tmp, "src")
        shutil.copytree(src, copy)
        shutil.rmtree(dst, ignore_errors=True)
        subprocess.check_call([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps",
                               "--find-links", "/opt/wheelhouse", "--target", dst, copy])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return dst


if __name__ == "__main__":
    print(install(*(sys.argv[1:2])))
