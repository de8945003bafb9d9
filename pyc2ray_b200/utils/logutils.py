def printlog(s, filename, quiet=False, end="\n"):
    """Append ``s`` to ``filename`` and echo it unless quiet (reference: pyc2ray/utils/logutils.py:1-15)."""
    if filename is not None:
        with open(filename, "a") as f:
            f.write(s + end)
    if not quiet:
        print(s, end=end)
