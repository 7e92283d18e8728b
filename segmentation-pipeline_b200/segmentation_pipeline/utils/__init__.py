"""Host-side helpers the hot path uses -- mirror of the reference's utils/utils.py (no_op :15-16,
is_sequence :19-20, as_list :23-28, collate_subjects :75-85, auto_str) and utils/config.py (Config :26-62).
The reference's TorchContext / TorchTimer / dataset tooling are outside the hot path and are not rebuilt."""
from .config import Config, get_nested_config
from .utils import no_op, is_sequence, as_list, as_tuple, collate_subjects, auto_str
