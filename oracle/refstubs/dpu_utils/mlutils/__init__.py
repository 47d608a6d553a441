"""TEST INFRASTRUCTURE ONLY -- minimal stand-in for `dpu_utils.mlutils` (absent wheel, dpu-utils>=0.2.17, setup.py:20).
`Vocabulary` implements the small token<->id surface the reference's embedders and task heads call when a model is built from
metadata in the tests (create_vocabulary / get_id_or_unk[_multiple] / is_unk / get_unk / get_pad / get_name_for_id / len); it is
a functional restatement of a frequency-ordered vocabulary, not a copy of dpu_utils.  None of it is arithmetic on the hot path."""
from collections import Counter
from typing import Iterable, List


class Vocabulary:
    _UNK, _PAD = "%UNK%", "%PAD%"

    def __init__(self, add_unk: bool = True, add_pad: bool = False):
        self._tokens: List[str] = []
        self._ids = {}
        if add_unk:
            self._add(self._UNK)
        if add_pad:
            self._add(self._PAD)

    def _add(self, tok: str) -> None:
        if tok not in self._ids:
            self._ids[tok] = len(self._tokens)
            self._tokens.append(tok)

    @classmethod
    def get_unk(cls) -> str:
        return cls._UNK

    @classmethod
    def get_pad(cls) -> str:
        return cls._PAD

    @classmethod
    def create_vocabulary(cls, tokens, max_size: int, count_threshold: int = 5, add_unk: bool = True, add_pad: bool = False):
        counts = tokens if isinstance(tokens, Counter) else Counter(tokens)
        vocab = cls(add_unk=add_unk, add_pad=add_pad)
        for tok, c in counts.most_common(max_size):
            if c < count_threshold or len(vocab) >= max_size:
                break
            vocab._add(tok)
        return vocab

    def __len__(self) -> int:
        return len(self._tokens)

    def is_unk(self, token: str) -> bool:
        return token not in self._ids

    def get_id_or_unk(self, token: str) -> int:
        return self._ids.get(token, self._ids.get(self._UNK, 0))

    def get_id_or_unk_multiple(self, tokens: Iterable[str], pad_to_size=None, padding_element: int = 0) -> List[int]:
        ids = [self.get_id_or_unk(t) for t in tokens]
        if pad_to_size is not None:
            ids = ids[:pad_to_size] + [padding_element] * max(0, pad_to_size - len(ids))
        return ids

    def get_name_for_id(self, token_id: int) -> str:
        return self._tokens[token_id]


class BpeVocabulary:  # pragma: no cover - import-time symbol only
    pass


class CharTensorizer:  # pragma: no cover - import-time symbol only
    pass
