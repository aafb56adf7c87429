"""Small host utilities behind the drop-in API: the logger whose error() is the
reference's error convention (mc3/utils/log.py:219-240 -> raises ValueError
with the message), burn-in masking (mc3/utils/utils.py:260-344) and default
parameter names (utils.py:347-361)."""
import sys
import textwrap
import time

import numpy as np


class Log:
    """stdout (+ optional file) logger with the reference's verbosity levels:
    verb < 0 errors only, >= 0 warnings, >= 1 head, >= 2 msg, >= 3 debug."""

    def __init__(self, logname=None, verb=2, append=False, width=70):
        self.logname = logname
        self.file = open(logname, 'a' if append else 'w') if logname else None
        self.verb, self.width, self.indent = verb, width, 0
        self.warnings = []
        self.sep = 70*':'

    def write(self, text):
        print(text)
        sys.stdout.flush()
        if self.file is not None and not self.file.closed:
            self.file.write(text + '\n')
            self.file.flush()

    def wrap(self, message, indent=None, si=None, width=None):
        indent = self.indent if indent is None else indent
        si = self.indent if si is None else si
        width = self.width if width is None else width
        lines = [textwrap.fill(s, width=width, initial_indent=' '*indent,
                               subsequent_indent=' '*si, break_long_words=False,
                               break_on_hyphens=False)
                 for s in message.splitlines()]
        return '\n'.join(lines)

    def _emit(self, level, message, **kw):
        if self.verb >= level:
            self.write(self.wrap(message, **kw))

    def msg(self, message, verb=2, **kw):
        self._emit(verb, message, **kw)

    def head(self, message, **kw):
        self._emit(1, message, **kw)

    def debug(self, message, **kw):
        self._emit(3, message, **kw)

    def warning(self, message):
        text = f'\n{self.sep}\n  Warning:\n{self.wrap(message, indent=4)}\n{self.sep}\n'
        self.warnings.append(message)
        if self.verb >= 0:
            self.write(text)

    def error(self, message, exception=ValueError):
        text = f'\n{self.sep}\n  Error:\n{self.wrap(message, indent=4)}\n{self.sep}\n'
        if self.file is not None and not self.file.closed:
            self.file.write(text + '\n')
            self.file.close()
        raise exception(message)

    def progressbar(self, frac):
        if self.verb >= 2:
            n = int(np.clip(10*frac, 0, 10))
            bar = ':'*n + ' '*(10 - n)
            self.write(f'\n[{bar}] {100*frac:5.1f}% completed  ({time.ctime()})')

    def close(self):
        if self.file is not None and not self.file.closed:
            self.file.close()


def tex_parameters(values, low_bounds, high_bounds, names=None, significant_digits=2):
    """LaTeX strings `value^{+hi}_{-lo}` with the number of decimals that shows
    `significant_digits` of the smaller error bar (mc3/utils/utils.py:364-468;
    a missing value prints the interval, a fixed parameter just the value)."""
    from decimal import Decimal
    out = []
    for k, value in enumerate(values):
        if value is None or np.isnan(value):
            lo, hi = low_bounds[k], high_bounds[k]
            place = Decimal(lo - hi).adjusted()
            dec = int(np.clip(significant_digits - 1 - place, 1, 10))
            text = f'[{lo:.{dec}f}, {hi:.{dec}f}]'
        else:
            lo, hi = low_bounds[k] - value, high_bounds[k] - value
            place = min(Decimal(lo).adjusted(), Decimal(hi).adjusted())
            dec = int(np.clip(significant_digits - 1 - place, 1, 10))
            text = f'{value}' if lo == hi else \
                f'{value:>.{dec}f}^{{{hi:+.{dec}f}}}_{{{lo:+.{dec}f}}}'
        prefix = '$'
        if names is not None:                       # names may carry their own math mode
            name = names[k].strip()
            prefix = f'{name[:-1]} = ' if name.startswith('$') and name.endswith('$') else f'{name}$ = '
        out.append(f'{prefix}{text}$')
    return out


def default_parnames(npars):
    return np.array([f'Param {i+1}' for i in range(npars)])


def burn(Zdict=None, burnin=None, Z=None, zchain=None, sort=True):
    """Drop the first `burnin` samples of every chain; returns (posterior,
    zchain, zmask) with the reference's ordering rules (utils.py:322-344)."""
    if Zdict is None and (Z is None or zchain is None or burnin is None):
        raise ValueError(
            'Need to input either Zdict or all three of burnin, Z, and zchain')
    if Zdict is not None:
        Z, zchain = Zdict['posterior'], Zdict['zchain']
        if burnin is None:
            burnin = Zdict['burnin']
    zchain = np.asarray(zchain)
    # Lock-step history (row k*nchains + c holds sample k of chain c): closed form.
    n = int(zchain.max()) + 1 if zchain.size else 0
    if n > 0 and zchain.size % n == 0 and zchain[0] == 0 and \
            np.array_equal(zchain.reshape(-1, n), np.broadcast_to(np.arange(n), (zchain.size//n, n))):
        K = zchain.size//n
        b = min(int(burnin), K)
        if sort:
            zmask = (np.arange(b, K)[None, :]*n + np.arange(n)[:, None]).ravel()
        else:
            zmask = np.arange(b*n, K*n)
        return Z[zmask], zchain[zmask], zmask
    # rank of each sample within its own chain, vectorised over chains
    order = np.argsort(zchain, kind='stable')
    zs = zchain[order]
    first = np.searchsorted(zs, zs, side='left')
    rank = np.empty(zchain.size, dtype=np.int64)
    rank[order] = np.arange(zchain.size) - first
    mask = (zchain >= 0) & (rank >= burnin)
    if sort:
        zmask = order[mask[order]]
    else:
        zmask = np.where(mask)[0]
    return Z[zmask], zchain[zmask], zmask
