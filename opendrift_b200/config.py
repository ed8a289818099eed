"""Configuration store of a simulation -- same keys, validation rules and error behaviour as the
reference's Configurable (opendrift/config.py:11-212), so that user scripts' set_config/get_config calls
are drop-in."""
import difflib
import logging

logger = logging.getLogger('opendrift_b200')

CONFIG_LEVEL_ESSENTIAL = 1
CONFIG_LEVEL_BASIC = 2
CONFIG_LEVEL_ADVANCED = 3

_REQUIRED = {'float': ('min', 'max', 'units'), 'int': ('min', 'max', 'units'),
             'str': ('min_length', 'max_length'), 'enum': ('enum',), 'bool': ()}


class Configurable:
    def __init__(self):
        self._config = {}

    # -- reading ----------------------------------------------------------------------------
    def get_config(self, key, default='raise'):
        item = self._config.get(key)
        if item is None:
            if default == 'raise':
                raise ValueError('No config setting named %s' % key)
            return default
        return item['value']

    def get_configspec(self, prefix='', level=(1, 2, 3)):
        levels = list(level) if isinstance(level, (list, tuple)) else [level]
        return {k: v for k, v in self._config.items() if k.startswith(prefix) and v['level'] in levels}

    def list_configspec(self, prefix=''):
        """config.py:34-52 -- one printed line per setting: name, value, type, range, start of the description."""
        for c, i in self._config.items():
            if not c.startswith(prefix):
                continue
            val = i.get('value')
            if i['type'] in ('bool',):
                rang = ''
            elif i['type'] == 'str':
                rang = 'min length %s, max length %s' % (i['min_length'], i['max_length'])
            elif i['type'] in ('float', 'int'):
                rang = 'min: %s, max: %s [%s]' % (i['min'], i['max'], i.get('units'))
            else:
                rang = i['enum']
            print('%-35s [%s] %-5s %s %s...' % (c, val, i['type'], rang, i['description'][0:20]))

    def list_config(self, prefix=''):
        lines = ['%s [%s]' % (k, v['value']) for k, v in self._config.items() if k.startswith(prefix)]
        logger.info('\n'.join(lines))
        return lines

    # -- writing ----------------------------------------------------------------------------
    def set_config(self, key, value):
        if isinstance(value, dict):                 # {'sub': v} -> key:sub
            for sub, v in value.items():
                self.set_config('%s:%s' % (key, sub), v)
            return
        if key not in self._config:
            raise ValueError('No config setting named %s' % key)
        spec = self._config[key]
        kind = spec['type']
        if kind == 'bool':
            if value not in (True, False):
                raise ValueError('Config value %s must be True or False' % key)
        elif kind in ('float', 'int'):
            if value is not None:
                lo, hi = spec['min'], spec['max']
                if (lo is not None and value < lo) or (hi is not None and value > hi):
                    raise ValueError('Config value %s must be between %s and %s' % (key, lo, hi))
                value = float(value) if kind == 'float' else int(value)
        elif kind == 'str':
            if not spec['min_length'] <= len(value) <= spec['max_length']:
                raise ValueError('String %s length must be between %s and %s characters'
                                 % (key, spec['min_length'], spec['max_length']))
        elif kind == 'enum':
            if value not in spec['enum']:
                hint = ''
                if len(spec['enum']) > 5 and isinstance(value, str):
                    low = {str(e).lower(): e for e in spec['enum']}
                    close = set(difflib.get_close_matches(value.lower(), list(low), n=20, cutoff=.3))
                    close |= {e for e in low if value.lower() in e}
                    if close:
                        hint = '\nDid you mean any of these?\n%s' % sorted(low[c] for c in close)
                raise ValueError('Wrong configuration (%s=%s), possible values are:\n\t%s\n%s'
                                 % (key, value, spec['enum'], hint))
        spec['value'] = value

    def _set_config_default(self, key, value):
        self.set_config(key, value)
        self._config[key]['default'] = self.get_config(key)

    def _add_config(self, config, overwrite=True):
        accepted = {}
        for key, spec in config.items():
            if key in self._config and not overwrite:
                continue
            for field in ('type', 'description', 'level'):
                if field not in spec:
                    raise ValueError('"%s" must be specified for config item %s' % (field, key))
            if spec['type'] not in _REQUIRED:
                raise ValueError('Config type "%s" (%s) is not defined. Valid options are: '
                                 'float, int, str, enum, bool' % (spec['type'], key))
            if spec['level'] != CONFIG_LEVEL_ESSENTIAL and 'default' not in spec:
                raise ValueError('A default value must be provided for config item %s' % key)
            for field in _REQUIRED[spec['type']]:
                if field not in spec:
                    raise ValueError('"%s" not provided for config item %s' % (field, key))
            if spec['type'] == 'enum' and not isinstance(spec['enum'], list):
                raise ValueError('"enum" of type list must be provided for config item %s' % key)
            spec = dict(spec)
            if 'default' in spec:
                spec['value'] = spec['default']
            spec.setdefault('value', None)
            accepted[key] = spec
        self._config.update(accepted)
