import sys
sys.path.insert(0,'/root/repo')
from dysgu_b200 import _lib
_lib.LIB_PATH = sys.argv[1]
sys.argv = [sys.argv[0]] + sys.argv[2:]
exec(open('/root/repo/scratch/stage_probe.py').read())
