"""User-visible error text of the DSL (same wording and layout as qbot/errors.py): a header
line, a window of up to five program lines with the failing one marked ``>>>``, then
``sys.exit()`` with no status."""
import sys

WINDOW = 5


def formatError(lines, lineNum, errorName, errorInfo):
    first = max(int(lineNum - (WINDOW - 1) / 2), 0)
    last = min(first + WINDOW, len(lines))
    width = len(str(last - 1))
    out = [f"{errorName}: {errorInfo}"]
    for i in range(first, last):
        mark = ">>> " if i == lineNum else "    "
        out.append(f"{mark}{str(i).zfill(width)}: {lines[i]}")
    return "\n".join(out)


def raiseFormattedError(error: str):
    print(error)
    sys.exit()


def customUnknownOperationError(lines, lineNum, op):
    return formatError(lines, lineNum, "UnknownOperation", op)


def customInvalidVariableName(lines, lineNum, name):
    return formatError(lines, lineNum, "InvalidVariableName", name)


def customInvalidMarkName(lines, lineNum, name):
    return formatError(lines, lineNum, "InvalidMarkName", name)


def customUnknownMarkName(lines, lineNum, name):
    return formatError(lines, lineNum, "UnknownMarkName", name)


def customNumArgumentsError(lines, lineNum, op, given, lo, hi=-1):
    if hi < lo:
        return formatError(lines, lineNum, "NumArgumentsError", f"operation {op} requires {lo}-{hi} arguments ({given} given)")
    return formatError(lines, lineNum, "NumArgumentsError", f"operation {op} requires {lo} argument(s) ({given} given)")


def customIndexError(lines, lineNum, what, index, maxIndex, minIndex=0):
    return formatError(lines, lineNum, "IndexError", f"{what} index {index} outside of valid range [{minIndex}, {maxIndex}]")


def customControlTargetOverlapError(lines, lineNum, index, lo, hi):
    if lo == hi:
        return formatError(lines, lineNum, "IndexError", f"control index {index} overlaps with target index {lo}")
    return formatError(lines, lineNum, "IndexError", f"control index {index} overlaps with target indices [{lo}, {hi}]")


def customTypeError(lines, lineNum, expected, got):
    exp = f"any of {expected}" if len(expected) > 1 else f"{expected[0]}"
    return formatError(lines, lineNum, "TypeError", f"{got} cannot be interpreted as {exp}")


def pythonError(lines, lineNum, e: Exception):
    return formatError(lines, lineNum, e.__class__.__name__, str(e))
