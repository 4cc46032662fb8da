#pragma once
#define LUSTRINE_EXPORT
