"""The C++ drop-in containers (include/clover_b200/containers.hpp): the reference README example and a
reference-style mvm validation compile against them (CPU) and run correctly on the GPU."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = "/tmp/clover_b200_readme_example"


def _build():
    import clover_b200
    clover_b200.build()
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++17", "-Wall", "-Werror", f"-I{ROOT}/include", f"{ROOT}/examples/readme_example.cpp",
           f"-L{ROOT}/clover_b200", "-lclover_b200", f"-Wl,-rpath,{ROOT}/clover_b200", "-o", EXE]
    subprocess.run(cmd, check=True, capture_output=True, text=True)


def test_example_compiles_against_dropin_containers():
    _build()
    assert os.path.exists(EXE)


@pytest.mark.gpu
def test_example_runs_on_gpu():
    _build()
    out = subprocess.run([EXE], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "The dot product is: 256" in out.stdout and out.stdout.strip().endswith("OK")
