"""TEST INFRASTRUCTURE — one rank of a `--devices N` command line on the CPU (tests/test_mgpu_cpu.py starts N of these with
PG_MG_RANK / PG_MG_WORLD / PG_MG_DIR set, which is how genomics_general_b200/mgpu.py recognises a rank process): the
engine is replaced by the oracle-backed stand-in, everything else is the product's multi-GPU path (the ranks are not children
of rank 0 here, as under torchrun).
Usage: python _mgpu_cpu_worker.py <cli module name> <argv ...>"""
import importlib
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

from oracle_engine_mg import OracleEngineMG  # noqa: E402

mod = importlib.import_module("genomics_general_b200.cli." + sys.argv[1])
mod.Engine = OracleEngineMG
mod.main(sys.argv[2:])
