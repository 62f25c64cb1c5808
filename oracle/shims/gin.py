"""Import shim for gin-config (absent here, no network): @gin.configurable becomes the identity.
Test infrastructure only -- lets the read-only reference under /root/reference import."""


def configurable(*args, **kwargs):
    if len(args) == 1 and callable(args[0]) and not kwargs:
        return args[0]

    def deco(fn):
        return fn
    return deco


def parse_config_files_and_bindings(*args, **kwargs):
    return None


REQUIRED = object()
