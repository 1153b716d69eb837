class ConfigItem:
    def __init__(self, defaultvalue='', description='', **kwargs):
        self.value = defaultvalue

    def __get__(self, obj, objtype=None):
        return self.value

    def __set__(self, obj, value):
        self.value = value


class ConfigNamespace:
    pass
