"""prego_b200 -- B200-native MiniROAD online step recognition + frame->step aggregation.

Drop-in for the hot path of aleflabo/PREGO (``step_recognition/model/rnn/rnn.py`` MROAD,
``trainer/eval.py`` Evaluate, ``utils/aggregate.py``) backed by hand-written sm_100a CUDA
behind the C ABI in ``include/prego_b200.h``.
"""
from .registry import META_ARCHITECTURES, EVAL, Registry, build_model, build_eval  # noqa: F401
from .model import MROAD, MROADA, FEATURE_SIZES  # noqa: F401
from .aggregate import aggregate, aggregate_labels  # noqa: F401
from .evaluate import Evaluate, ANT_Evaluate  # noqa: F401
from .metrics import perframe_average_precision  # noqa: F401
from .pipeline import predict_labels, recognize_and_aggregate  # noqa: F401
from .training import (OadLoss, build_criterion, train_one_step, allreduce_gradients, enable_overlapped_allreduce, FusedAdamW,  # noqa: F401
                       TRAINER, train_one_epoch, build_trainer, build_optimizer, WindowDataset, gradient_buckets)
from . import ingest  # noqa: F401

__version__ = "0.1.0"
