"""Drop-in module `sg_net`: put this directory on PYTHONPATH (or copy the reference scripts beside it) and
`from sg_net import ...` in eval_pair.py / eval_batch.py / main_sg.py resolves to the B200-native package."""
from sg_pr_b200.sg_net import *  # noqa: F401,F403
from sg_pr_b200.sg_net import SG, SGTrainer  # noqa: F401
