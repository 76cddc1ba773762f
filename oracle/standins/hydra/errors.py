class InstantiationException(Exception):
    pass
