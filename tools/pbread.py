"""CLI: dump a frozen GraphDef.  `python tools/pbread.py graph.pb [verbose]`."""
import os, sys, runpy
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from sentinel_tree_cover_b200.pbread import *  # noqa: F401,F403
if __name__ == "__main__":
    runpy.run_module("sentinel_tree_cover_b200.pbread", run_name="__main__")
