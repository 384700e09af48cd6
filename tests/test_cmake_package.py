"""The build-system surface of the drop-in boundary: a consumer that does
`find_package(pumitally)` + `target_link_libraries(app pumitally::pumitally)` (what the OpenMC
fork does with the reference, reference CMakeLists.txt:113-138) configures and links against
this repo's CMake package without CUDA or NCCL on its own command line."""
import os
import shutil
import subprocess

import pytest

from helpers import ROOT

cmake = shutil.which("cmake")


@pytest.mark.skipif(cmake is None or shutil.which("ninja") is None, reason="cmake/ninja not available")
def test_find_package_consumer(tmp_path):
    build, inst, app = tmp_path / "build", tmp_path / "inst", tmp_path / "app"
    build.mkdir(); app.mkdir()
    common = ["-G", "Ninja", "-DCMAKE_CXX_COMPILER=/usr/bin/g++"]
    subprocess.check_call([cmake, *common, "-DCMAKE_CUDA_COMPILER=/usr/local/cuda/bin/nvcc",
                           "-DCMAKE_CUDA_HOST_COMPILER=/usr/bin/g++", f"-DCMAKE_INSTALL_PREFIX={inst}", ROOT],
                          cwd=build, stdout=subprocess.DEVNULL)
    subprocess.check_call(["ninja", "install"], cwd=build, stdout=subprocess.DEVNULL)
    assert (inst / "include" / "pumitally" / "PumiTally.h").exists()
    assert (inst / "lib" / "cmake" / "pumitally" / "pumitallyConfig.cmake").exists()
    (app / "CMakeLists.txt").write_text(
        "cmake_minimum_required(VERSION 3.20)\nproject(app CXX)\n"
        "find_package(pumitally REQUIRED)\n"
        f"add_executable(app {os.path.join(ROOT, 'tests', 'facade_demo.cpp')})\n"
        "target_compile_features(app PRIVATE cxx_std_17)\n"
        "target_link_libraries(app PRIVATE pumitally::pumitally)\n")
    appb = tmp_path / "appbuild"
    appb.mkdir()
    subprocess.check_call([cmake, *common, f"-DCMAKE_PREFIX_PATH={inst}", str(app)], cwd=appb, stdout=subprocess.DEVNULL)
    subprocess.check_call(["ninja"], cwd=appb, stdout=subprocess.DEVNULL)
    assert (appb / "app").exists()
    needed = subprocess.check_output(["readelf", "-d", str(appb / "app")], text=True)
    assert "libpumitally.so" in needed and "libcudart" not in needed and "libnccl" not in needed
