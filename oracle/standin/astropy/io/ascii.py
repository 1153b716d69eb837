from ..table import Table


def read(filename, **kwargs):
    return Table.read(filename, format=kwargs.pop('format', 'ascii'))
