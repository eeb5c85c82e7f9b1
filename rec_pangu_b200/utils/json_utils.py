import json


def beautify_json(ori_dict: dict) -> str:
    """Formatted JSON string (reference: rec_pangu/utils/json_utils.py:12-22; syntax colouring when pygments exists)."""
    s = json.dumps(ori_dict, indent=4, ensure_ascii=False, sort_keys=True)
    try:
        from pygments import highlight, formatters, lexers
        return highlight(s, lexers.get_lexer_by_name('json'), formatters.TerminalFormatter())
    except Exception:
        return s
