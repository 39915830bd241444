"""Stand-in for ftfy (lib/models/simple_tokenizer.py:30): names only."""
def fix_text(t):
    return t
