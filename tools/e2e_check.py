import os, sys, runpy
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
runpy.run_path(os.path.join(ROOT, "tests", "e2echeck.py"), run_name="__main__")
