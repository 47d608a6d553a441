def split_identifier_into_parts(identifier):  # pragma: no cover - import-time symbol only
    return [identifier]
