"""TEST INFRASTRUCTURE ONLY -- import-time stand-in for ``dpu-utils`` (`/root/reference/setup.py:20`).
Only the symbols the reference imports at module load are provided; none is arithmetic on the
message-passing path."""
