"""Stand-in for open_clip (lib/models/utils.py:9): names only."""
def create_model_from_pretrained(*a, **k):
    raise NotImplementedError("open_clip stub")
def get_tokenizer(*a, **k):
    raise NotImplementedError("open_clip stub")
