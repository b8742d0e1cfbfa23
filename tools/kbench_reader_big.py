import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.kbench_reader import run
run("nuscenes", 8, 2_000_000)
