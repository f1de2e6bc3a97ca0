"""Drop-in module `dgcnn`: put this directory on PYTHONPATH (or copy the reference scripts beside it) and
`from dgcnn import ...` in eval_pair.py / eval_batch.py / main_sg.py resolves to the B200-native package."""
from sg_pr_b200.dgcnn import *  # noqa: F401,F403
