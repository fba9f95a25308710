"""ecog2txt_b200 -- B200-native engine for the ECoG->text seq2seq hot path of jgmakin/ecog2txt.

Boundary constants follow /root/reference/ecog2txt/__init__.py:13-22.
"""
EOS_token = '<EOS>'
pad_token = '<pad>'
OOV_token = '<OOV>'
TOKEN_TYPES = {'phoneme', 'word', 'trial', 'word_sequence', 'word_piece_sequence', 'phoneme_sequence'}
DATA_PARTITIONS = {'training', 'validation', 'testing'}

from .engine import Engine, EngineConfig, E2TError  # noqa: E402,F401
from .sequence_network import SequenceNetwork  # noqa: E402,F401
from .trainers import MultiSubjectTrainer  # noqa: E402,F401
