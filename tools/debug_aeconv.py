import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import parity as P
for case in [(64, 2, 2, 96, 256, 1, 1), (64, 2, 2, 32, 96, 1, 1), (64, 1, 1, 200, 128, 1, 1), (64, 1, 1, 128, 200, 1, 1),
             (64, 4, 4, 256, 256, 5, 2), (64, 7, 7, 128, 256, 5, 2), (64, 14, 14, 64, 128, 5, 2), (64, 28, 28, 1, 64, 5, 2),
             (16, 2, 2, 256, 96, 1, 1), (16, 4, 4, 256, 256, 5, 2), (16, 8, 8, 128, 256, 5, 2), (16, 16, 16, 64, 128, 5, 2)]:
    print(case, P.conv_case(*case), P.conv_case(*case, with_mask=True)["dgrad"], flush=True)
