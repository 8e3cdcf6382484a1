"""assistants — Python host-side mirror of the reference's ReAct loop, the CALLER of the Chat seam
(reference pkg/assistants/simple.go:292-616 `AssistantWithConfig`).  Same control flow, same strings:

  first Chat -> json.Unmarshal into ToolPrompt; a reply that is not JSON is returned verbatim         (simple.go:343-382)
  loop (<= maxIterations, default 5):                                                                 (simple.go:391-412)
      final_answer set, not a template placeholder, and an observation present -> return it           (simple.go:414-419)
      action.name set -> run tools[name](input); errors/unknown tools become the observation text     (simple.go:421-481)
      observation = ConstrictPrompt(observation, model, 1024); whole ToolPrompt marshalled as a USER message (simple.go:495-501)
      Chat again; reply with final_answer -> return it; unparsable reply -> "Summarize all the chat history…" Chat (simple.go:515-600)

Used by bench/tests to drive multi-step tool-calling loops through the engine; `tools` is injectable because the reference's
pkg/tools shell out to kubectl/trivy (out of scope: SURVEY.md §2 #6)."""
from __future__ import annotations

import json
from dataclasses import dataclass, field
from .perf import GetPerfStats

from .llms import ChatCompletionMessage, ChatMessageRoleAssistant, ChatMessageRoleUser, ConstrictPrompt, TrimSpace

defaultMaxIterations = 5                                     # simple.go:22
_TEMPLATE_PATTERNS = ["<最终答案", "<final_answer", "<Final answer", "<最终回答", "<回答", "<答案", "使用 Markdown 格式", "使用Markdown格式",
                      "换行符用 \\n 表示", "换行符用\\n表示"]


def isTemplateValue(value: str) -> bool:                     # simple.go:624-657
    if len(value.encode("utf-8")) < 10:
        return True
    if any(p in value for p in _TEMPLATE_PATTERNS):
        return True
    return "<" in value and ">" in value


@dataclass
class ToolPrompt:                                            # reference pkg/tools/tool.go:29-38
    question: str = ""
    thought: str = ""
    action: dict = field(default_factory=lambda: {"name": "", "input": ""})
    observation: str = ""
    final_answer: str = ""

    @classmethod
    def unmarshal(cls, text: str) -> "ToolPrompt":
        """json.Unmarshal([]byte(text), &toolPrompt) (simple.go:366) as encoding/json does it: the document's keys are visited IN ORDER, each one
        matched to a field exactly or else case-insensitively (so duplicates and differently-cased duplicates overwrite each other, the last one in the
        document wins); unknown keys are ignored; JSON null leaves the field as it is; any other non-string value for a string field, or a non-object
        `action`, is an UnmarshalTypeError — decoding goes on, but Unmarshal returns that error, which sends the caller into its 'not JSON, assume final
        answer' / 'Summarize…' branches.  Nothing is coerced."""
        class Obj(list):
            pass
        d = json.loads(text, object_pairs_hook=Obj)
        if not isinstance(d, Obj):
            raise ValueError("json: cannot unmarshal non-object into Go value of type tools.ToolPrompt")
        tp = cls()
        errors = []

        def fold(k: str) -> str:                     # encoding/json's foldName (simple case folding): U+017F folds to 's', U+212A (Kelvin) to 'k' — str.lower() does the latter
            return k.replace("\u017f", "s").lower()

        def set_string(target, attr, v, path):
            if v is None:
                return
            if not isinstance(v, str):
                errors.append(f"json: cannot unmarshal {type(v).__name__} into Go struct field ToolPrompt.{path} of type string")
                return
            if isinstance(target, dict):
                target[attr] = v
            else:
                setattr(target, attr, v)
        for k, v in d:
            f = fold(k)
            if f in ("question", "thought", "observation", "final_answer"):
                set_string(tp, f, v, f)
            elif f == "action":
                if v is None:
                    continue
                if not isinstance(v, Obj):
                    errors.append("json: cannot unmarshal non-object into Go struct field ToolPrompt.action")
                    continue
                for k2, v2 in v:
                    f2 = fold(k2)
                    if f2 in ("name", "input"):
                        set_string(tp.action, f2, v2, "action." + f2)
        if errors:
            raise ValueError(errors[0])
        return tp

    def marshal(self) -> str:
        r"""json.Marshal(toolPrompt) (simple.go:497), byte for byte: struct field order, no spaces, and Go's default HTML-safe string
        escaping — '<', '>', '&' become \u003c / \u003e / \u0026, U+2028 / U+2029 become \u2028 / \u2029, other non-ASCII stays
        raw UTF-8, control bytes use \n \r \t or \u00XX.  kubectl's ubiquitous "<none>" would otherwise tokenise differently from
        what the Go caller sends."""
        return ("{" + f'"question":{go_json_string(self.question)},"thought":{go_json_string(self.thought)},'
                f'"action":{{"name":{go_json_string(self.action["name"])},"input":{go_json_string(self.action["input"])}}},'
                f'"observation":{go_json_string(self.observation)},"final_answer":{go_json_string(self.final_answer)}' + "}")


def go_json_string(s: str) -> str:
    """encoding/json's string encoder with escapeHTML=true (the json.Marshal default)."""
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif ch == "\b":
            out.append("\\b")            # Go >= 1.22 writes \b and \f in short form (the reference builds with go 1.24: go.mod:3, Dockerfile:2)
        elif ch == "\f":
            out.append("\\f")
        elif o < 0x20 or ch in "<>&" or o in (0x2028, 0x2029):
            out.append("\\u%04x" % o)
        elif 0xD800 <= o <= 0xDFFF:
            out.append("\\ufffd")          # invalid UTF-8 in a Go string marshals as U+FFFD
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def AssistantWithConfig(model, prompts, maxTokens, countTokens, verbose, maxIterations, client, tools, count_tokens=None):
    """-> (result, chatHistory).  `client` has Chat(model, maxTokens, prompts); `tools` maps name -> callable(input) -> str."""
    perf = GetPerfStats()
    total_done = perf.TraceFunc("assistant_total")                                          # simple.go:296
    try:
        return _assistant_loop(perf, model, prompts, maxTokens, maxIterations, client, tools, count_tokens)
    finally:
        total_done()


def _timed_chat(perf, op, client, model, maxTokens, chatHistory):
    perf.StartTimer(op)
    try:
        return client.Chat(model, maxTokens, chatHistory)
    except Exception as e:
        raise RuntimeError(f"chat completion error: {e}") from e
    finally:
        perf.StopTimer(op)


def _assistant_loop(perf, model, prompts, maxTokens, maxIterations, client, tools, count_tokens):
    chatHistory = list(prompts)
    if not prompts:
        raise ValueError("prompts cannot be empty")                                          # simple.go:312
    resp = _timed_chat(perf, "assistant_first_chat", client, model, maxTokens, chatHistory)  # simple.go:341-346
    chatHistory.append(ChatCompletionMessage(ChatMessageRoleAssistant, resp))
    perf.StartTimer("assistant_parse_tool_prompt")                                           # simple.go:364-385
    try:
        tp = ToolPrompt.unmarshal(resp)
    except Exception:
        return resp, chatHistory                                                             # not JSON: assume final answer
    finally:
        perf.StopTimer("assistant_parse_tool_prompt")
    iterations = 0
    if maxIterations <= 0:
        maxIterations = defaultMaxIterations
    while True:
        iterations += 1
        if iterations > maxIterations:
            return tp.final_answer, chatHistory
        if tp.final_answer != "" and not isTemplateValue(tp.final_answer) and tp.observation != "":
            return tp.final_answer, chatHistory
        if tp.action["name"] != "":
            fn = tools.get(tp.action["name"])
            if fn is not None:
                perf.StartTimer("assistant_tool_" + tp.action["name"])                        # simple.go:440-475
                try:
                    observation = TrimSpace(fn(tp.action["input"]))                             # strings.TrimSpace (simple.go:444)
                except Exception as e:
                    observation = f"Tool {tp.action['name']} failed with error {e}. Considering refine the inputs for the tool."
                finally:
                    perf.StopTimer("assistant_tool_" + tp.action["name"])
            else:
                observation = f"Tool {tp.action['name']} is not available. Considering switch to other supported tools."
            perf.StartTimer("assistant_construct_message")                                   # simple.go:491-507
            tp.observation = ConstrictPrompt(observation, model, 1024, count_tokens)
            chatHistory.append(ChatCompletionMessage(ChatMessageRoleUser, tp.marshal()))
            perf.StopTimer("assistant_construct_message")
            resp = _timed_chat(perf, "assistant_intermediate_chat", client, model, maxTokens, chatHistory)   # simple.go:513-518
            chatHistory.append(ChatCompletionMessage(ChatMessageRoleAssistant, resp))
            perf.StartTimer("assistant_parse_intermediate")                                  # simple.go:541-603
            try:
                tp = ToolPrompt.unmarshal(resp)
                perf.StopTimer("assistant_parse_intermediate")
            except Exception:
                perf.StopTimer("assistant_parse_intermediate")
                chatHistory.append(ChatCompletionMessage(ChatMessageRoleUser,
                                                         "Summarize all the chat history and respond to original question with final answer"))
                resp = _timed_chat(perf, "assistant_summarize", client, model, maxTokens, chatHistory)        # simple.go:564-569
                return resp, chatHistory
            if tp.final_answer != "":
                return tp.final_answer, chatHistory
