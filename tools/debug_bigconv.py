import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P
for case in [(512, 16, 16, 200, 400, 5, 2), (512, 8, 8, 400, 800, 5, 2), (512, 32, 32, 3, 200, 5, 2), (512, 1, 1, 200, 12800, 1, 1)]:
    print(case, P.conv_case(*case), flush=True)
