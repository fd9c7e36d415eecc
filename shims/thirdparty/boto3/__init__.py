"""boto3 stub: reference file_utils.py:32 imports it for S3 model downloads, which need network access (out of scope)."""
