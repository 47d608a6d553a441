class Vocabulary:  # pragma: no cover - import-time symbol only
    pass


class BpeVocabulary:  # pragma: no cover
    pass


class CharTensorizer:  # pragma: no cover
    pass
