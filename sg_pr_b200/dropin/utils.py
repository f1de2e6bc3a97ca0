"""Drop-in module `utils`: put this directory on PYTHONPATH (or copy the reference scripts beside it) and
`from utils import ...` in eval_pair.py / eval_batch.py / main_sg.py resolves to the B200-native package."""
from sg_pr_b200.utils import *  # noqa: F401,F403
from sg_pr_b200.utils import json, math, np, os, random  # noqa: F401  (the reference's star-import hands these out)
