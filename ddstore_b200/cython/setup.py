"""Builds the `pyddstore` Cython extension in-tree (ddstore_b200/cython/pyddstore*.so), linked against
../libddstore_b200.so. Run from this directory: python setup.py build_ext --inplace"""
import os

from Cython.Build import cythonize
from setuptools import Extension, setup

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
ROOT = os.path.dirname(PKG)

ext = Extension(
    "pyddstore",
    sources=[os.path.join(HERE, "pyddstore.pyx")],
    language="c++",
    include_dirs=[os.path.join(ROOT, "include")],
    libraries=["ddstore_b200"],
    library_dirs=[PKG],
    runtime_library_dirs=["$ORIGIN/.."],
    extra_compile_args=["-std=c++17", "-O2"],
)
setup(name="pyddstore", ext_modules=cythonize([ext], language_level=3, quiet=True), script_args=None)
