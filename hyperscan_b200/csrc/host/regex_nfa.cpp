/* regex_nfa.cpp -- see regex_nfa.h */
#include "regex_nfa.h"

#include <algorithm>
#include <bitset>
#include <cstring>
#include <memory>

#include "../../../include/hs_b200.h"

namespace hsb {

namespace {

typedef std::bitset<256> CharSet;

struct Node {
    enum Kind { CLASS, CAT, ALT, STAR, PLUS, OPT, EMPTY, ASSERT } kind = EMPTY;
    int assertKind = 0;    /* ASSERT: one of the A_* bits below */
    CharSet cls;
    std::vector<std::shared_ptr<Node>> kids;
};
/* zero-width assertions; a path through several of them carries the union of their bits */
enum {
    A_WORD_B = 1,      /* \\b */
    A_NOT_WORD_B = 2,  /* \\B */
    A_BEGIN = 4,       /* \\A, "^": offset 0 */
    A_BEGIN_LINE = 8,  /* "^" under HS_FLAG_MULTILINE: offset 0 or right after a newline */
    A_END = 16,        /* \\z: end of the data */
    A_END_LF = 32,     /* "$", \\Z: end of the data, or before a newline that ends the data */
    A_END_LINE = 64,   /* "$" under HS_FLAG_MULTILINE: end of the data or before any newline */
    A_BOUNDARY = A_WORD_B | A_NOT_WORD_B,
    A_STARTS = A_BEGIN | A_BEGIN_LINE,
    A_ENDS = A_END | A_END_LF | A_END_LINE,
    A_ALL = 127
};
typedef std::shared_ptr<Node> NodeP;

NodeP mk(Node::Kind k) {
    NodeP n = std::make_shared<Node>();
    n->kind = k;
    return n;
}

NodeP clone(const NodeP &n) {
    NodeP c = std::make_shared<Node>(*n);
    for (NodeP &k : c->kids) {
        k = clone(k);
    }
    return c;
}

class Parser {
public:
    Parser(const char *re_in, unsigned flags_in) : re(re_in), n(strlen(re_in)), flags(flags_in) {}

    NodeP parse() {
        if (flags & (HS_FLAG_UTF8 | HS_FLAG_UCP)) {
            fail("HS_FLAG_UTF8 / HS_FLAG_UCP need the reference's Unicode compiler.");
        }
        if (flags & HS_FLAG_SOM_LEFTMOST) {
            fail("HS_FLAG_SOM_LEFTMOST is not supported.");
        }
        NodeP r = alt();
        if (pos != n) {
            fail("Unmatched parentheses.");
        }
        /* anchors only where the reference takes them (ComponentBoundary::checkEmbeddedStartAnchor / EndAnchor,
         * src/parser/ComponentBoundary.cpp:162-185, and the callers' rules in ComponentSequence / Alternation / Repeat) */
        embeddedStart(r, true);
        embeddedEnd(r, true);
        return r;
    }

private:
    const char *re;
    size_t n, pos = 0;
    unsigned flags;
    size_t unrolled = 0;  /* positions made by unrolling bounded repeats so far */
    bool quoting = false; /* between \\Q and \\E: every byte is a literal */

    [[noreturn]] void fail(const std::string &m) const { throw RegexError{m}; }
    bool at(char c) const { return pos < n && re[pos] == c; }

    CharSet fold(CharSet s) const {
        if (flags & HS_FLAG_CASELESS) {
            for (int c = 'a'; c <= 'z'; c++) {
                if (s[c] || s[c - 32]) {
                    s[c] = s[c - 32] = true;
                }
            }
        }
        return s;
    }

    NodeP alt() {
        std::vector<NodeP> arms;
        arms.push_back(seq());
        while (at('|')) {
            pos++;
            arms.push_back(seq());
        }
        if (arms.size() == 1) {
            return arms[0];
        }
        NodeP a = mk(Node::ALT);
        a->kids = arms;
        return a;
    }

    NodeP seq() {
        NodeP s = mk(Node::CAT);
        while (pos < n && (quoting || (re[pos] != '|' && re[pos] != ')'))) {
            NodeP a = atom();
            a = quantified(a);
            s->kids.push_back(a);
        }
        return s;
    }

    /* may the match still start here (nothing consumed before on any way to this node)?  A start anchor
     * anywhere else is refused, as the reference does. */
    static bool embeddedStart(const NodeP &x, bool atStart) {
        switch (x->kind) {
        case Node::CLASS:
            return false;
        case Node::EMPTY:
            return atStart;
        case Node::ASSERT:
            if (x->assertKind & A_BOUNDARY) {
                return false;
            }
            if ((x->assertKind & A_STARTS) && !atStart) {
                throw RegexError{"Embedded start anchors not supported."};
            }
            return atStart;
        case Node::CAT:
            for (const NodeP &k : x->kids) {
                atStart = embeddedStart(k, atStart);
            }
            return atStart;
        case Node::ALT: {
            bool rv = true;
            for (const NodeP &k : x->kids) {
                rv &= embeddedStart(k, atStart);
            }
            return rv;
        }
        case Node::STAR:
        case Node::PLUS:
            atStart = embeddedStart(x->kids[0], atStart);
            return embeddedStart(x->kids[0], atStart);
        case Node::OPT:
            return embeddedStart(x->kids[0], atStart);
        }
        return false;
    }
    static bool embeddedEnd(const NodeP &x, bool atEnd) {
        switch (x->kind) {
        case Node::CLASS:
            return false;
        case Node::EMPTY:
            return atEnd;
        case Node::ASSERT:
            if (x->assertKind & A_BOUNDARY) {
                return false;
            }
            if ((x->assertKind & A_ENDS) && !atEnd) {
                throw RegexError{"Embedded end anchors not supported."};
            }
            return atEnd;
        case Node::CAT:
            for (auto k = x->kids.rbegin(); k != x->kids.rend(); ++k) {
                atEnd = embeddedEnd(*k, atEnd);
            }
            return atEnd;
        case Node::ALT: {
            bool rv = true;
            for (const NodeP &k : x->kids) {
                rv &= embeddedEnd(k, atEnd);
            }
            return rv;
        }
        case Node::STAR:
        case Node::PLUS:
            atEnd = embeddedEnd(x->kids[0], atEnd);
            return embeddedEnd(x->kids[0], atEnd);
        case Node::OPT:
            return embeddedEnd(x->kids[0], atEnd);
        }
        return false;
    }

    static size_t leaves(const NodeP &x) {
        size_t c = x->kind == Node::CLASS || x->kind == Node::ASSERT;
        for (const NodeP &k : x->kids) {
            c += leaves(k);
        }
        return c;
    }

    NodeP repeat(const NodeP &a, unsigned lo, long hi /* -1 = unbounded */) {
        /* repeats are unrolled: refuse before copying what cannot fit the largest model anyway */
        unrolled += leaves(a) * std::max<size_t>(hi < 0 ? lo + 1 : (size_t)hi, 1);
        if (unrolled > 4 * MAX_NFA_STATES) {
            fail("Pattern is too large.");
        }
        NodeP s = mk(Node::CAT);
        for (unsigned i = 0; i < lo; i++) {
            s->kids.push_back(clone(a));
        }
        if (hi < 0) {
            if (lo == 0) {
                NodeP st = mk(Node::STAR);
                st->kids.push_back(clone(a));
                s->kids.push_back(st);
            } else { /* x{n,} = x^(n-1) x+ */
                NodeP pl = mk(Node::PLUS);
                pl->kids.push_back(s->kids.back());
                s->kids.back() = pl;
            }
        } else {
            for (long i = lo; i < hi; i++) { /* x{n,m}: m - n optional copies (same language as the nested form) */
                NodeP o = mk(Node::OPT);
                o->kids.push_back(clone(a));
                s->kids.push_back(o);
            }
        }
        return s;
    }

    /* "{n}", "{n,}", "{n,m}" at q?  (anything else that starts with a brace is literal text) */
    bool startsRepeat(size_t q) const {
        if (q >= n || re[q] != '{') {
            return false;
        }
        size_t k = q + 1;
        while (k < n && isdigit((unsigned char)re[k])) k++;
        if (k == q + 1) {
            return false;
        }
        if (k < n && re[k] == ',') {
            k++;
            while (k < n && isdigit((unsigned char)re[k])) k++;
        }
        return k < n && re[k] == '}';
    }

    NodeP quantified(NodeP a) {
        while (pos + 1 < n && re[pos] == '\\' && re[pos + 1] == 'E') { /* "\\Qab\\E+": the quantifier binds to the "b" */
            quoting = false;
            pos += 2;
        }
        if (pos >= n || quoting) {
            return a;
        }
        const char c = re[pos];
        if (a->kind == Node::ASSERT && (c == '*' || c == '+' || c == '?' || c == '{')) {
            fail("A quantifier on an assertion is not supported.");
        }
        unsigned lo = 1;
        long hi = 1;
        if (c == '*') {
            pos++;
            lo = 0;
            hi = -1;
        } else if (c == '+') {
            pos++;
            lo = 1;
            hi = -1;
        } else if (c == '?') {
            pos++;
            lo = 0;
            hi = 1;
        } else if (c == '{') {
            size_t q = pos + 1;
            unsigned long x = 0, y = 0;
            bool haveX = false, haveY = false, comma = false;
            while (q < n && isdigit((unsigned char)re[q])) {
                x = std::min(x * 10 + (unsigned long)(re[q++] - '0'), 100000ul);
                haveX = true;
            }
            if (q < n && re[q] == ',') {
                comma = true;
                q++;
                while (q < n && isdigit((unsigned char)re[q])) {
                    y = std::min(y * 10 + (unsigned long)(re[q++] - '0'), 100000ul);
                    haveY = true;
                }
            }
            if (!haveX || q >= n || re[q] != '}') {
                return a; /* not a quantifier: the brace is a literal (atom() reads it next) */
            }
            if (comma && haveY && y < x) {
                fail("Bounded repeat is invalid: min > max.");
            }
            if (x > 1000 || (haveY && y > 1000)) {
                fail("Pattern is too large.");
            }
            pos = q + 1;
            lo = (unsigned)x;
            hi = !comma ? (long)x : haveY ? (long)y : -1;
        } else {
            return a;
        }
        if (pos < n && re[pos] == '?') {
            pos++; /* lazy: same set of match ends */
        } else if (pos < n && re[pos] == '+') {
            fail("Possessive quantifiers are not supported.");
        }
        if (lo == 1 && hi == 1) {
            return a;
        }
        if (lo == 0 && hi == 0) {
            return mk(Node::EMPTY);
        }
        return repeat(a, lo, hi);
    }

    static int hexval(char c) {
        if (c >= '0' && c <= '9') return c - '0';
        if (c >= 'a' && c <= 'f') return c - 'a' + 10;
        if (c >= 'A' && c <= 'F') return c - 'A' + 10;
        return -1;
    }

    static CharSet classEscape(char e) {
        CharSet s;
        switch (e | 0x20) {
        case 'd':
            for (int c = '0'; c <= '9'; c++) s[c] = true;
            break;
        case 'w':
            for (int c = '0'; c <= '9'; c++) s[c] = true;
            for (int c = 'a'; c <= 'z'; c++) s[c] = s[c - 32] = true;
            s['_'] = true;
            break;
        case 's': /* PCRE 8.41 \s: space, \t \n \v \f \r */
            s[' '] = s['\t'] = s['\n'] = s[0x0b] = s['\f'] = s['\r'] = true;
            break;
        }
        return (e >= 'A' && e <= 'Z') ? ~s : s;
    }

    static bool posixClass(const std::string &name, CharSet *out) { /* C locale, as PCRE */
        static const struct { const char *n; int (*f)(int); } tab[] = {
            {"alpha", isalpha}, {"digit", isdigit}, {"alnum", isalnum}, {"upper", isupper}, {"lower", islower},
            {"space", isspace}, {"punct", ispunct}, {"print", isprint}, {"graph", isgraph}, {"cntrl", iscntrl},
            {"xdigit", isxdigit}, {"blank", isblank}};
        for (const auto &t : tab) {
            if (name == t.n) {
                for (int c = 0; c < 128; c++) {
                    if (t.f(c)) (*out)[(size_t)c] = true;
                }
                return true;
            }
        }
        if (name == "word") {
            *out = classEscape('w');
            return true;
        }
        if (name == "ascii") {
            for (int c = 0; c < 128; c++) (*out)[(size_t)c] = true;
            return true;
        }
        return false;
    }

    /* an escape: either one character (*single) or a class */
    CharSet escape(bool inClass, int *single) {
        *single = -1;
        if (++pos >= n) {
            fail("Unterminated escape at end of pattern.");
        }
        const unsigned char e = (unsigned char)re[pos++];
        CharSet s;
        switch (e) {
        case 'n': *single = '\n'; break;
        case 't': *single = '\t'; break;
        case 'r': *single = '\r'; break;
        case 'f': *single = '\f'; break;
        case 'a': *single = '\a'; break;
        case 'e': *single = 0x1b; break;
        case 'd': case 'D': case 'w': case 'W': case 's': case 'S':
            return classEscape((char)e);
        case 'h': case 'H': /* horizontal white space, 8-bit: \t, space, 0xa0 */
            s['\t'] = s[' '] = s[0xa0] = true;
            return e == 'H' ? ~s : s;
        case 'v': case 'V': /* vertical white space, 8-bit: \n \v \f \r, 0x85 */
            s['\n'] = s[0x0b] = s['\f'] = s['\r'] = s[0x85] = true;
            return e == 'V' ? ~s : s;
        case 'N': /* not a newline, whatever HS_FLAG_DOTALL says */
            s.set();
            s['\n'] = false;
            return s;
        case 'x': {
            if (pos < n && re[pos] == '{') { /* \x{h..h} */
                size_t q = pos + 1;
                unsigned long v = 0;
                while (q < n && hexval(re[q]) >= 0) {
                    v = std::min(v * 16 + (unsigned long)hexval(re[q]), 0x10000ul);
                    q++;
                }
                if (q == pos + 1 || q >= n || re[q] != '}') {
                    fail("Invalid hex escape: \\x{ without hexadecimal digits and a closing brace.");
                }
                if (v > 0xff) {
                    fail("Hexadecimal value is greater than \\xFF.");
                }
                pos = q + 1;
                *single = (int)v;
                break;
            }
            int v = 0; /* \x followed by up to two hexadecimal digits (none: a NUL byte) */
            for (int k = 0; k < 2 && pos < n && hexval(re[pos]) >= 0; k++) {
                v = v * 16 + hexval(re[pos++]);
            }
            *single = v;
            break;
        }
        case 'c': { /* control character: the next byte, upper-cased, with bit 6 flipped */
            if (pos >= n) {
                fail("\\c at end of pattern.");
            }
            const unsigned char x = (unsigned char)re[pos++];
            if (x >= 128) {
                fail("\\c must be followed by an ASCII character.");
            }
            *single = toupper(x) ^ 0x40;
            break;
        }
        case '0': { /* \0 and up to two more octal digits */
            int v = 0;
            for (int k = 0; k < 2 && pos < n && re[pos] >= '0' && re[pos] <= '7'; k++) {
                v = v * 8 + (re[pos++] - '0');
            }
            *single = v;
            break;
        }
        default:
            if (inClass && e == 'b') {
                *single = 0x08;
            } else if (inClass && e >= '1' && e <= '7') { /* no back-references in a class: octal, up to three digits */
                int v = e - '0';
                for (int k = 0; k < 2 && pos < n && re[pos] >= '0' && re[pos] <= '7'; k++) {
                    v = v * 8 + (re[pos++] - '0');
                }
                *single = v & 0xff;
            } else if (isalnum(e)) {
                fail(std::string("Escape sequence \\") + (char)e + " is not supported.");
            } else {
                *single = e;
            }
        }
        s[(size_t)*single] = true;
        return s;
    }

    CharSet charClass() {
        pos++; /* [ */
        bool negate = false;
        if (at('^')) {
            negate = true;
            pos++;
        }
        CharSet set;
        bool first = true;   /* no member yet: a "]" is a literal */
        bool quoted = false; /* between \\Q and \\E */
        int prev = -1;
        for (;;) {
            if (pos >= n) {
                fail("Unterminated character class.");
            }
            const unsigned char c = (unsigned char)re[pos];
            if (c == '\\' && pos + 1 < n && (re[pos + 1] == 'E' || (re[pos + 1] == 'Q' && !quoted))) {
                quoted = re[pos + 1] == 'Q';
                pos += 2;
                continue;
            }
            if (quoted) {
                set[c] = true;
                prev = c;
                pos++;
                first = false;
                continue;
            }
            if (c == ']' && !first) {
                pos++;
                break;
            }
            first = false;
            if (c == '[' && pos + 1 < n && re[pos + 1] == ':') {
                const char *close = strstr(re + pos + 2, ":]");
                if (!close) {
                    fail("Unterminated POSIX character class.");
                }
                std::string name(re + pos + 2, close);
                bool neg = false;
                if (!name.empty() && name[0] == '^') {
                    neg = true;
                    name.erase(0, 1);
                }
                CharSet pc;
                if (!posixClass(name, &pc)) {
                    fail("Unknown POSIX character class.");
                }
                pc = fold(pc); /* under CASELESS [:upper:] and [:lower:] are all letters, THEN the "^" applies */
                set |= neg ? ~pc : pc;
                pos = (size_t)(close - re) + 2;
                prev = -1;
                continue;
            }
            if (c == '[' && pos + 1 < n && (re[pos + 1] == '.' || re[pos + 1] == '=')) {
                fail("POSIX collating elements are not supported.");
            }
            if (c == '-' && prev >= 0 && pos + 1 < n && re[pos + 1] != ']') {
                pos++;
                int hi;
                if (re[pos] == '[' && pos + 1 < n && strchr(":.=", re[pos + 1])) {
                    fail("Invalid range in character class.");
                }
                if (re[pos] == '\\') {
                    if (pos + 1 < n && (re[pos + 1] == 'Q' || re[pos + 1] == 'E')) {
                        fail("\\Q / \\E as the end of a range is not supported.");
                    }
                    escape(true, &hi);
                    if (hi < 0) {
                        fail("Invalid range in character class.");
                    }
                } else {
                    hi = (unsigned char)re[pos++];
                }
                if (hi < prev) {
                    fail("Invalid range in character class.");
                }
                for (int v = prev; v <= hi; v++) {
                    set[(size_t)v] = true;
                }
                prev = -1;
                continue;
            }
            if (c == '\\') {
                int single;
                const CharSet e = escape(true, &single);
                set |= e;
                prev = single;
            } else {
                set[c] = true;
                prev = c;
                pos++;
            }
        }
        set = fold(set);
        return negate ? ~set : set;
    }

    NodeP leaf(const CharSet &s) {
        if (s.none()) {
            fail("Empty character class.");
        }
        NodeP l = mk(Node::CLASS);
        l->cls = s;
        return l;
    }

    NodeP atom() {
        /* \\Q ... \\E: literal bytes (an \\E without \\Q is ignored) */
        while (pos + 1 < n && re[pos] == '\\' && (re[pos + 1] == 'E' || (re[pos + 1] == 'Q' && !quoting))) {
            quoting = re[pos + 1] == 'Q';
            pos += 2;
        }
        if (pos >= n || (!quoting && (re[pos] == '|' || re[pos] == ')'))) {
            return mk(Node::EMPTY);
        }
        const unsigned char c = (unsigned char)re[pos];
        if (quoting) {
            pos++;
            CharSet s;
            s[c] = true;
            return leaf(fold(s));
        }
        if (c == '(') {
            pos++;
            const unsigned saved = flags; /* an option change lasts to the end of the group it is in */
            if (at('?')) {
                /* (?ims-ims) and (?ims-ims: ... ) */
                size_t q = pos + 1;
                unsigned on = 0, off = 0;
                bool neg = false, any = false;
                for (; q < n && strchr("ims-", re[q]); q++) {
                    if (re[q] == '-') {
                        neg = true;
                        continue;
                    }
                    const unsigned f = re[q] == 'i' ? HS_FLAG_CASELESS : re[q] == 's' ? HS_FLAG_DOTALL : HS_FLAG_MULTILINE;
                    (neg ? off : on) |= f;
                    any = true;
                }
                if (q < n && re[q] == ')' && any) {
                    flags = (flags | on) & ~off;
                    pos = q + 1;
                    if (pos < n && (strchr("*+?", re[pos]) || startsRepeat(pos))) {
                        fail("Invalid repeat.");
                    }
                    return mk(Node::EMPTY);
                }
                if (q < n && re[q] == ':') {
                    flags = (flags | on) & ~off;
                    pos = q + 1;
                } else {
                    fail("Look-around, named groups, comments and the other (?...) groups are not supported.");
                }
            }
            NodeP r = alt();
            if (!at(')')) {
                fail("Missing close parenthesis.");
            }
            pos++;
            flags = saved;
            return r;
        }
        if (c == '[') {
            /* "[:name:]", "[.x.]", "[=x=]" where a class should start: PCRE's check_posix_syntax */
            if (pos + 1 < n && strchr(":.=", re[pos + 1])) {
                const char term = re[pos + 1];
                for (size_t q = pos + 2; q < n; q++) {
                    if (re[q] == '\\' && q + 1 < n && (re[q + 1] == ']' || re[q + 1] == '\\')) {
                        q++;
                    } else if ((re[q] == '[' && q + 1 < n && re[q + 1] == term) || re[q] == ']') {
                        break;
                    } else if (re[q] == term && q + 1 < n && re[q + 1] == ']') {
                        fail(term == ':' ? "POSIX named classes are only supported inside a class."
                                         : "Unsupported POSIX collating element.");
                    }
                }
            }
            return leaf(charClass());
        }
        if (c == '\\' && pos + 1 < n && strchr("bBAzZ", re[pos + 1])) {
            NodeP a = mk(Node::ASSERT);
            const char e = re[pos + 1];
            a->assertKind = e == 'b' ? A_WORD_B : e == 'B' ? A_NOT_WORD_B : e == 'A' ? A_BEGIN : e == 'z' ? A_END : A_END_LF;
            pos += 2;
            return a;
        }
        if (c == '^' || c == '$') {
            NodeP a = mk(Node::ASSERT);
            const bool ml = flags & HS_FLAG_MULTILINE;
            a->assertKind = c == '^' ? (ml ? A_BEGIN_LINE : A_BEGIN) : (ml ? A_END_LINE : A_END_LF);
            pos++;
            return a;
        }
        if (c == '\\') {
            int single;
            CharSet s = escape(false, &single);
            return leaf(single >= 0 ? fold(s) : s);
        }
        if (c == '.') {
            pos++;
            CharSet s;
            s.set();
            if (!(flags & HS_FLAG_DOTALL)) {
                s['\n'] = false;
            }
            return leaf(s);
        }
        if (strchr("*+?", c) || (c == '{' && startsRepeat(pos))) {
            fail("Invalid repeat: nothing to repeat.");
        }
        pos++;
        CharSet s;
        s[c] = true;
        return leaf(fold(s));
    }
};

/* Glushkov: positions, nullable / first / last / follow */
struct Glushkov {
    std::vector<CharSet> cls;         /* per position */
    std::vector<StateSet> follow;     /* per position: set of positions */
    std::vector<int> assertion;       /* per position: 0 = a character position, else the A_* bit of a zero-width
                                       * pseudo-position (eliminated when the NFA is wired) */
    struct Sets {
        bool nullable;
        StateSet first, last;
    };
    static const u32 MAX_POSITIONS = MAX_NFA_STATES - 2; /* the two start states come first */

    Sets build(const NodeP &n) {
        switch (n->kind) {
        case Node::EMPTY:
            return {true, StateSet(), StateSet()};
        case Node::CLASS: {
            if (cls.size() >= MAX_POSITIONS) {
                throw RegexError{"Pattern is too large: more than 510 character positions."};
            }
            const u32 p = (u32)cls.size();
            cls.push_back(n->cls);
            follow.push_back(StateSet());
            assertion.push_back(0);
            return {false, stateBit(p), stateBit(p)};
        }
        case Node::ASSERT: {
            if (cls.size() >= MAX_POSITIONS) {
                throw RegexError{"Pattern is too large: more than 510 character positions."};
            }
            const u32 p = (u32)cls.size();
            cls.push_back(CharSet());
            follow.push_back(StateSet());
            assertion.push_back(n->assertKind);
            return {false, stateBit(p), stateBit(p)};
        }
        case Node::CAT: {
            Sets acc = {true, StateSet(), StateSet()};
            for (const NodeP &k : n->kids) {
                const Sets s = build(k);
                link(acc.last, s.first);
                const StateSet first = acc.nullable ? (acc.first | s.first) : acc.first;
                const StateSet last = s.nullable ? (acc.last | s.last) : s.last;
                acc = {acc.nullable && s.nullable, first, last};
            }
            return acc;
        }
        case Node::ALT: {
            Sets acc = {false, StateSet(), StateSet()};
            for (const NodeP &k : n->kids) {
                const Sets s = build(k);
                acc = {acc.nullable || s.nullable, acc.first | s.first, acc.last | s.last};
            }
            return acc;
        }
        case Node::STAR:
        case Node::PLUS: {
            const Sets s = build(n->kids[0]);
            link(s.last, s.first);
            return {n->kind == Node::STAR || s.nullable, s.first, s.last};
        }
        case Node::OPT: {
            const Sets s = build(n->kids[0]);
            return {true, s.first, s.last};
        }
        }
        return {true, StateSet(), StateSet()};
    }

    void link(const StateSet &from, const StateSet &to) {
        for (u32 p = 0; p < follow.size(); p++) {
            if (from.test(p)) {
                follow[p] |= to;
            }
        }
    }

    /* ---- zero-width assertions --------------------------------------------------------------
     * A path between two character positions p -> q may cross \b / \B pseudo-positions; what they
     * assert is about the two characters on either side: (p is a word character) != / == (q is).
     * need: bit 0 = some \b crossed, bit 1 = some \B crossed. */
    static bool holds(u32 need, bool leftWord, bool rightWord) {
        return !((need & 1) && leftWord == rightWord) && !((need & 2) && leftWord != rightWord);
    }
    bool isWord(u32 p) const { /* positions of an expression with assertions are split: all word or none */
        static const CharSet W = wordSet();
        return (cls[p] & W).any();
    }
    static CharSet wordSet() {
        CharSet w;
        for (int c = '0'; c <= '9'; c++) w[(size_t)c] = true;
        for (int c = 'a'; c <= 'z'; c++) w[(size_t)c] = w[(size_t)(c - 32)] = true;
        w['_'] = true;
        return w;
    }
    /* character positions reachable from the position set `from` through assertion pseudo-positions only,
     * each with the assertions crossed: out[(q, need)] */
    void reachThroughAssertions(const StateSet &from, u32 need, std::vector<std::pair<u32, u32>> *out,
                                std::vector<u8> *seen /* [position][need] */) const {
        for (u32 x = 0; x < cls.size(); x++) {
            if (!from.test(x)) {
                continue;
            }
            if (!assertion[x]) {
                out->push_back({x, need});
                continue;
            }
            const u32 n2 = need | (u32)assertion[x];
            u8 &mark = (*seen)[x * (A_ALL + 1) + n2];
            if (mark) {
                continue;
            }
            mark = 1;
            reachThroughAssertions(follow[x], n2, out, seen);
        }
    }
    /* the assertion sets under which a path from p (exclusive) through assertions only ends in `last` */
    void exitsThroughAssertions(const StateSet &from, u32 need, const StateSet &last, std::vector<u32> *needs, std::vector<u8> *seen) const {
        for (u32 x = 0; x < cls.size(); x++) {
            if (!from.test(x) || !assertion[x]) {
                continue;
            }
            const u32 n2 = need | (u32)assertion[x];
            u8 &mark = (*seen)[x * (A_ALL + 1) + n2];
            if (mark) {
                continue;
            }
            mark = 1;
            if (last.test(x)) {
                needs->push_back(n2);
            }
            exitsThroughAssertions(follow[x], n2, last, needs, seen);
        }
    }
    /* the assertion sets of the ways from `first` to `last` that cross assertions only (no character) */
    void emptyPaths(const StateSet &from, u32 need, const StateSet &last, std::vector<u32> *needs, std::vector<u8> *seen) const {
        for (u32 x = 0; x < cls.size(); x++) {
            if (!from.test(x) || !assertion[x]) {
                continue;
            }
            const u32 n2 = need | (u32)assertion[x];
            u8 &mark = (*seen)[x * (A_ALL + 1) + n2];
            if (mark) {
                continue;
            }
            mark = 1;
            if (last.test(x)) {
                needs->push_back(n2);
            }
            emptyPaths(follow[x], n2, last, needs, seen);
        }
    }
    bool anyAssertion() const {
        for (int a : assertion) {
            if (a) return true;
        }
        return false;
    }
};

bool hasAssertion(const NodeP &n, int kinds) {
    if (n->kind == Node::ASSERT && (n->assertKind & kinds)) {
        return true;
    }
    for (const NodeP &k : n->kids) {
        if (hasAssertion(k, kinds)) return true;
    }
    return false;
}

/* an expression with \b / \B: every class that mixes word and non-word characters becomes an
 * alternation of its two halves, so that each character position is one or the other */
void splitByWordness(NodeP &n) {
    if (n->kind == Node::CLASS) {
        const CharSet W = Glushkov::wordSet();
        const CharSet w = n->cls & W, o = n->cls & ~W;
        if (w.any() && o.any()) {
            NodeP a = mk(Node::ALT);
            NodeP l = mk(Node::CLASS), r = mk(Node::CLASS);
            l->cls = w;
            r->cls = o;
            a->kids = {l, r};
            n = a;
        }
        return;
    }
    for (NodeP &k : n->kids) {
        splitByWordness(k);
    }
}

u32 minLenOf(const NodeP &n) {
    switch (n->kind) {
    case Node::EMPTY: return 0;
    case Node::CLASS: return 1;
    case Node::CAT: {
        u32 s = 0;
        for (const NodeP &k : n->kids) s += minLenOf(k);
        return s;
    }
    case Node::ALT: {
        u32 m = ~0u;
        for (const NodeP &k : n->kids) m = std::min(m, minLenOf(k));
        return m;
    }
    case Node::PLUS: return minLenOf(n->kids[0]);
    case Node::STAR:
    case Node::OPT: return 0;
    }
    return 0;
}

u64 maxLenOf(const NodeP &n) { /* > 0xfffffffe = unbounded */
    const u64 INF = 1ull << 40;
    switch (n->kind) {
    case Node::EMPTY: return 0;
    case Node::CLASS: return 1;
    case Node::CAT: {
        u64 s = 0;
        for (const NodeP &k : n->kids) s = std::min(INF, s + maxLenOf(k));
        return s;
    }
    case Node::ALT: {
        u64 m = 0;
        for (const NodeP &k : n->kids) m = std::max(m, maxLenOf(k));
        return m;
    }
    case Node::PLUS:
    case Node::STAR: return maxLenOf(n->kids[0]) ? INF : 0;
    case Node::OPT: return maxLenOf(n->kids[0]);
    }
    return 0;
}

/* top-level alternatives of the expression, each with its anchored flag */
std::vector<NodeP> topArms(const NodeP &root) {
    NodeP r = root; /* "(a|b)" as a whole expression is the alternation a|b */
    while (r->kind == Node::CAT && r->kids.size() == 1) {
        r = r->kids[0];
    }
    if (r->kind == Node::ALT) {
        return r->kids;
    }
    return {root};
}

} // namespace

RegexInfo regexInfo(const char *re, unsigned flags, bool forInfo) {
    const NodeP root = Parser(re, flags).parse();
    RegexInfo info;
    info.minLen = ~0u;
    bool someOtherExit = false, someFloatingEntry = false;
    /* one way out of the expression: the assertions crossed after the last character, and whether that character
     * is a word character (-1: there is none, the byte before the match is whatever it is) */
    auto exitVia = [&](u32 need, int leftWord) {
        info.unordered |= (need & (A_BOUNDARY | A_END_LF | A_END_LINE)) != 0;
        bool canEod = false;
        if (need & (A_ENDS | A_BOUNDARY)) {
            for (int lw = 0; lw < 2; lw++) {
                if (leftWord >= 0 && lw != leftWord) {
                    continue;
                }
                if ((need & A_STARTS) && lw) {
                    continue; /* the start of the data / a newline before: not a word character */
                }
                canEod |= Glushkov::holds(need, lw != 0, false);
            }
        }
        info.atEod |= canEod;
        if (!(need & (A_END | A_END_LF))) {
            someOtherExit = true;
        }
    };
    for (NodeP arm : topArms(root)) {
        const bool asserts = hasAssertion(arm, A_BOUNDARY);
        if (asserts) {
            splitByWordness(arm);
        }
        /* an assertion at the end may look one byte ahead */
        info.needsAdjust |= hasAssertion(arm, A_BOUNDARY | A_END_LF | A_END_LINE);
        Glushkov g;
        const Glushkov::Sets s = g.build(arm);
        if (!forInfo && (s.nullable || minLenOf(arm) == 0)) {
            throw RegexError{"Pattern matches empty buffer; use HS_FLAG_ALLOWEMPTY to enable support."};
        }
        info.minLen = std::min(info.minLen, minLenOf(arm));
        const u64 mx = maxLenOf(arm);
        info.maxLen = std::max<u32>(info.maxLen, mx > 0xfffffffeull ? 0xffffffffu : (u32)mx);
        info.armWidths.push_back({minLenOf(arm), mx > 0xfffffffeull ? 0xffffffffu : (u32)mx});
        info.positions += (u32)g.cls.size();
        const u32 np = (u32)g.cls.size();
        for (u32 p = 0; p < np; p++) {
            if (g.assertion[p]) {
                continue;
            }
            const int lw = asserts ? (int)g.isWord(p) : -1;
            if (s.last.test(p)) {
                exitVia(0, lw);
            }
            std::vector<u32> needs;
            std::vector<u8> seen(np * (A_ALL + 1), 0);
            g.exitsThroughAssertions(g.follow[p], 0, s.last, &needs, &seen);
            for (u32 need : needs) {
                exitVia(need, lw);
            }
        }
        {
            std::vector<std::pair<u32, u32>> entries;
            std::vector<u8> seenE(np * (A_ALL + 1), 0);
            g.reachThroughAssertions(s.first, 0, &entries, &seenE);
            for (const auto &e : entries) {
                someFloatingEntry |= !(e.second & A_BEGIN);
            }
        }
        if (s.nullable) {
            exitVia(0, -1);
            someFloatingEntry = true;
        }
        std::vector<u32> needs;
        std::vector<u8> seen(np * (A_ALL + 1), 0);
        g.emptyPaths(s.first, 0, s.last, &needs, &seen);
        for (u32 need : needs) {
            exitVia(need, -1);
        }
    }
    info.onlyAtEod = info.atEod && !someOtherExit;
    info.anchored = !someFloatingEntry;
    return info;
}

void regexNfaInit(RawNfa *nfa) {
    *nfa = RawNfa();
    nfa->nstates = 2; /* 0 = floating start (.* loop), 1 = anchored start (offset 0 only) */
    nfa->succ.assign(2, StateSet());
    nfa->succ[0] = stateBit(0);
    nfa->squashMask.assign(2, allStates());
    nfa->squashKind.assign(2, LIMEX_SQUASH_NONE);
    nfa->reports.resize(2);
    nfa->reportsEod.resize(2);
    for (u32 b = 0; b < 256; b++) {
        nfa->reach[b] = stateBit(0); /* the floating start survives every byte; nothing re-enters state 1 */
    }
    nfa->init = nfa->initDS = stateSetOf(3);
}

void regexNfaAdd(RawNfa *nfa, const char *re, unsigned flags, u32 report, u32 reportBeforeNewline, u64 minLength) {
    const NodeP root = Parser(re, flags).parse();
    const CharSet W = Glushkov::wordSet();
    auto newState = [&]() -> u32 {
        if (nfa->nstates >= MAX_NFA_STATES) {
            throw RegexError{"Pattern set is too large: its character positions exceed the 512-state NFA model."};
        }
        nfa->nstates++;
        nfa->succ.push_back(StateSet());
        nfa->squashMask.push_back(allStates());
        nfa->squashKind.push_back(LIMEX_SQUASH_NONE);
        nfa->reports.emplace_back();
        nfa->reportsEod.emplace_back();
        return nfa->nstates - 1;
    };
    auto reachClass = [&](u32 st, const CharSet &c) {
        for (u32 b = 0; b < 256; b++) {
            if (c[b]) {
                nfa->reach[b].set(st);
            }
        }
    };
    for (NodeP arm : topArms(root)) {
        const bool asserts = hasAssertion(arm, A_BOUNDARY);
        if (asserts) {
            splitByWordness(arm);
        }
        Glushkov g;
        const Glushkov::Sets s = g.build(arm);
        if (s.nullable || minLenOf(arm) == 0) {
            throw RegexError{"Pattern matches empty buffer; use HS_FLAG_ALLOWEMPTY to enable support."};
        }
        const u32 np = (u32)g.cls.size();
        /* hs_expr_ext.min_length: only matches of at least that many bytes count.  The automaton then carries the
         * number of bytes consumed since the match began, saturating at the minimum: K copies ("levels") of every
         * position, a transition goes one level up, and only the top level accepts.  (No minimum, or every match
         * is long enough anyway: one level.) */
        if (minLength > MAX_NFA_STATES) {
            throw RegexError{"Pattern is too large: min_length beyond the 512-state NFA model."};
        }
        const u32 K = minLength > minLenOf(arm) ? (u32)minLength : 1;
        /* NFA states of every character position, one per level (assertion pseudo-positions get none) */
        std::vector<std::vector<u32>> lv(np, std::vector<u32>(K, 0));
        for (u32 p = 0; p < np; p++) {
            if (!g.assertion[p]) {
                for (u32 c = 0; c < K; c++) {
                    lv[p][c] = newState();
                    reachClass(lv[p][c], g.cls[p]);
                }
            }
        }
        /* the accepting copy of a position */
        std::vector<u32> st(np, 0);
        for (u32 p = 0; p < np; p++) {
            st[p] = lv[p][K - 1];
        }

        /* --- transitions between character positions --- */
        for (u32 p = 0; p < np; p++) {
            if (g.assertion[p]) {
                continue;
            }
            std::vector<std::pair<u32, u32>> to;
            std::vector<u8> seen(np * (A_ALL + 1), 0);
            g.reachThroughAssertions(g.follow[p], 0, &to, &seen);
            for (const auto &e : to) {
                if (e.second & (A_STARTS | A_ENDS)) {
                    continue; /* an anchor between two characters (the parser lets none through) */
                }
                if (Glushkov::holds(e.second, g.isWord(p), g.isWord(e.first))) {
                    for (u32 c = 0; c < K; c++) {
                        nfa->succ[lv[p][c]].set(lv[e.first][std::min(c + 1, K - 1)]);
                    }
                }
            }
        }

        /* --- entries --- */
        std::vector<std::pair<u32, u32>> entries;
        {
            std::vector<u8> seen(np * (A_ALL + 1), 0);
            g.reachThroughAssertions(s.first, 0, &entries, &seen);
        }
        for (const auto &e : entries) {
            const u32 entry = lv[e.first][0]; /* one byte consumed */
            const bool qw = asserts && g.isWord(e.first);
            if (e.second & A_ENDS) {
                continue; /* a character after the end of the data */
            }
            if (e.second & A_STARTS) {
                /* what precedes is the start of the data or a newline: not a word character */
                if (!Glushkov::holds(e.second, false, qw)) {
                    continue;
                }
                nfa->succ[1].set(entry);
                if (!(e.second & A_BEGIN)) {
                    /* "^" under (?m): entered at offset 0 (anchored start) or right after any newline --
                     * one state shared by all such entries, on after every '\n' */
                    if (!nfa->mlStartState) {
                        nfa->mlStartState = newState();
                        nfa->reach[(u8)'\n'].set(nfa->mlStartState);
                        nfa->succ[0].set(nfa->mlStartState);
                    }
                    nfa->succ[nfa->mlStartState].set(entry);
                }
            } else if (e.second == 0) {
                nfa->succ[0].set(entry);
            } else {
                /* a leading \b / \B of a floating entry: the byte before the match decides.  Two
                 * shared context states hang off the floating start: "previous byte is a word character"
                 * and "previous byte is none, or not a word character" (the latter also on at offset 0) */
                if (!nfa->ctxWord) {
                    nfa->ctxWord = newState();
                    nfa->ctxNonWord = newState();
                    reachClass(nfa->ctxWord, W);
                    reachClass(nfa->ctxNonWord, ~W);
                    nfa->succ[0].set(nfa->ctxWord).set(nfa->ctxNonWord);
                    nfa->init.set(nfa->ctxNonWord);
                    nfa->initDS.set(nfa->ctxNonWord);
                }
                if (Glushkov::holds(e.second, true, qw)) {
                    nfa->succ[nfa->ctxWord].set(entry);
                }
                if (Glushkov::holds(e.second, false, qw)) {
                    nfa->succ[nfa->ctxNonWord].set(entry);
                }
            }
        }

        /* --- accepts --- */
        auto adjusted = [&]() -> u32 {
            if (!reportBeforeNewline) {
                throw RegexError{"internal: no adjusted report program"};
            }
            return reportBeforeNewline;
        };
        auto addOnce = [](std::vector<u32> &v, u32 r) {
            if (std::find(v.begin(), v.end(), r) == v.end()) {
                v.push_back(r);
            }
        };
        /* one byte of look-ahead, reported one byte back (offset_adjust -1, as the reference compiles "$"):
         * a word character, a non-word character (trailing \b / \B), any newline ("$" under (?m)), a newline
         * that ends the data ("$", \Z) */
        u32 aheadWord = 0, aheadNonWord = 0, aheadNewline = 0, aheadLastNewline = 0;
        for (u32 p = 0; p < np; p++) {
            if (g.assertion[p]) {
                continue;
            }
            if (s.last.test(p)) {
                addOnce(nfa->reports[st[p]], report);
            }
            std::vector<u32> needs;
            std::vector<u8> seen(np * (A_ALL + 1), 0);
            g.exitsThroughAssertions(g.follow[p], 0, s.last, &needs, &seen);
            const bool pw = asserts && g.isWord(p);
            for (u32 need : needs) {
                if (need & A_STARTS) {
                    continue; /* a start anchor after a character (the parser lets none through) */
                }
                if (Glushkov::holds(need, pw, false)) { /* the end of the data counts as a non-word character */
                    addOnce(nfa->reportsEod[st[p]], report);
                }
                if (need & A_ENDS) {
                    if ((need & A_END) || !Glushkov::holds(need, pw, false)) {
                        continue; /* \z: nothing but the end of the data will do; a newline is not a word character */
                    }
                    if (need & A_END_LF) {
                        if (!aheadLastNewline) {
                            aheadLastNewline = newState();
                            nfa->reach[(u8)'\n'].set(aheadLastNewline);
                            nfa->reportsEod[aheadLastNewline].push_back(adjusted());
                        }
                        nfa->succ[st[p]].set(aheadLastNewline);
                    } else {
                        if (!aheadNewline) {
                            aheadNewline = newState();
                            nfa->reach[(u8)'\n'].set(aheadNewline);
                            nfa->reports[aheadNewline].push_back(adjusted());
                        }
                        nfa->succ[st[p]].set(aheadNewline);
                    }
                    continue;
                }
                if (Glushkov::holds(need, pw, true)) {
                    if (!aheadWord) {
                        aheadWord = newState();
                        reachClass(aheadWord, W);
                        nfa->reports[aheadWord].push_back(adjusted());
                    }
                    nfa->succ[st[p]].set(aheadWord);
                }
                if (Glushkov::holds(need, pw, false)) {
                    if (!aheadNonWord) {
                        aheadNonWord = newState();
                        reachClass(aheadNonWord, ~W);
                        nfa->reports[aheadNonWord].push_back(adjusted());
                    }
                    nfa->succ[st[p]].set(aheadNonWord);
                }
            }
        }
    }
}

} // namespace hsb
