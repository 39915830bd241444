"""Stand-in for OpenAI clip (lib/models/utils.py:18): names only."""
