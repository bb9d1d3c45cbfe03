import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.conv_check import run
a = [int(v) for v in sys.argv[1:]]
print(run(*a[:10], stats=bool(a[10]) if len(a) > 10 else True))
