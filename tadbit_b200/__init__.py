"""tadbit_b200 -- B200-native Hi-C matrix balancing (TADbit's normalize_hic hot path).

``HiC_data`` keeps the reference's Python surface; the cells live in HBM and every
reduction (filter statistics, ICE row sums and bias updates, normalised sum, per-diagonal
sums) is a hand-written sm_100a CUDA kernel behind the C-ABI of include/tadbit_b200.h.
There is no CPU fallback.
"""
from ._capi import B200Error, SynthParams, device_count
from .batch import batch_store, ice_batch, normalize_hic_batch
from .decay import column_biases, rescale_and_decay
from .experiment import Experiment
from . import extraction
from .hic_data import HiC_data, from_reference
from .hic_filtering import filter_by_mean, filter_by_zero_count
from .normalize_hic import expected, iterative
from .store import CommGroup, ContactStore
from .vectors import BinVector, PairTable

__version__ = "0.1.0"
__all__ = ["HiC_data", "from_reference", "ContactStore", "CommGroup", "SynthParams", "BinVector", "PairTable",
           "iterative", "expected", "filter_by_mean", "filter_by_zero_count",
           "device_count", "B200Error", "normalize_hic_batch", "batch_store", "ice_batch", "Experiment", "rescale_and_decay", "column_biases", "extraction"]
